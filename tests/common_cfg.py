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

# QuartzNet-style stack (tensor-core-aligned): separable convolutions everywhere -- a stride-2 first layer on the
# 64 features (composed dense kernel), depthwise + pointwise blocks with the reference's single (non-dense)
# residual whose 1x1 conv is separable too, a dilated layer, a kernel_size-1 separable layer and a plain 1x1 conv.
MINI_QUARTZ = [
    {"type": "sep_conv1d", "repeat": 1, "kernel_size": [11], "stride": [2], "num_channels": 256,
     "padding": "SAME", "dilation": [1]},
    {"type": "sep_conv1d", "repeat": 2, "kernel_size": [13], "stride": [1], "num_channels": 256,
     "padding": "SAME", "dilation": [1], "residual": True, "residual_dense": False},
    {"type": "sep_conv1d", "repeat": 2, "kernel_size": [33], "stride": [1], "num_channels": 384,
     "padding": "SAME", "dilation": [1], "residual": True, "residual_dense": False},
    {"type": "sep_conv1d", "repeat": 1, "kernel_size": [15], "stride": [1], "num_channels": 384,
     "padding": "SAME", "dilation": [2], "residual": True, "residual_dense": False},
    {"type": "sep_conv1d", "repeat": 1, "kernel_size": [1], "stride": [1], "num_channels": 256,
     "padding": "SAME", "dilation": [1]},
    {"type": "conv1d", "repeat": 1, "kernel_size": [1], "stride": [1], "num_channels": 256,
     "padding": "SAME", "dilation": [1]},
]
