"""t2v -- host-side glue of the B200-native Tacotron2-VAE hot path.

`_lib`    ctypes binding of libt2v_b200.so (C ABI in include/t2v_b200.h)
`build`   nvcc recipe (sm_100a) for the library
`engine`  forward/backward orchestration of the kernels behind model.Tacotron2
"""
