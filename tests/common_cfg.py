"""Small tensor-core-aligned Jasper-like encoder configs shared by the GPU parity tests."""

MINI_JASPER = [
    {"type": "conv1d", "repeat": 1, "kernel_size": [11], "stride": [2], "num_channels": 128,
     "padding": "SAME", "dilation": [1], "dropout_keep_prob": 0.8},
    {"type": "conv1d", "repeat": 2, "kernel_size": [5], "stride": [1], "num_channels": 128,
     "padding": "SAME", "dilation": [1], "dropout_keep_prob": 0.8, "residual": True, "residual_dense": True},
    {"type": "conv1d", "repeat": 2, "kernel_size": [7], "stride": [1], "num_channels": 256,
     "padding": "SAME", "dilation": [1], "dropout_keep_prob": 0.8, "residual": True, "residual_dense": True},
    {"type": "conv1d", "repeat": 1, "kernel_size": [9], "stride": [1], "num_channels": 256,
     "padding": "SAME", "dilation": [2], "dropout_keep_prob": 0.6},
    {"type": "conv1d", "repeat": 1, "kernel_size": [1], "stride": [1], "num_channels": 256,
     "padding": "SAME", "dilation": [1], "dropout_keep_prob": 0.6},
]
