"""dlgrad/runtime/cuda: the Device.CUDA runtime — cffi shim over libdlgrad_cuda.so (driver API +
NCCL), nvcc kernel cache, and the (op, Device.CUDA) functions registered with the dispatcher."""
