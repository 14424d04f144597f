def load_pre_existing_vocabulary(path, min_idx=0, read_chars=False):
    """open_seq2seq/data/utils.py:28-58: token -> index in file order."""
    idx = min_idx
    vocab = {}
    with open(path, "r", encoding="utf-8") as f:
        for line in f:
            if not line or line == "\n":
                continue
            if read_chars:
                token = line[0]
            else:
                token = line.rstrip().split("\t")[0]
            vocab[token] = idx
            idx += 1
    return vocab
