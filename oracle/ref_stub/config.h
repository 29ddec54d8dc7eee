/* oracle/_ref build: CPU only (BUILD_WITH_CUDA undefined), see oracle/ref_stub/tf_stub.h */
