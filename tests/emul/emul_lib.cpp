// emul_lib.cpp -- TEST INFRASTRUCTURE ONLY: the engine's sources compiled as C++ on top of
// cuda_shim.h, giving tests a CPU-run replica of the kernels' control logic with the same C ABI.
// Built into tests/emul/libsalt_b200_emul.so by tests/emul/build_emul.py; never shipped.
#include "cuda_shim.h"

namespace emu {
thread_local dim3 t_threadIdx, t_blockIdx, t_blockDim, t_gridDim;
thread_local CtaState *t_cta = nullptr;
}

#include "../../salt_b200/csrc/verify.cu"
#include "../../salt_b200/csrc/ssw.cu"
#include "../../salt_b200/csrc/samtail.cu"
#include "../../salt_b200/csrc/mixref.cu"
#include "../../salt_b200/csrc/transport.cu"
#include "../../salt_b200/csrc/seed.cu"
#include "../../salt_b200/csrc/engine.cu"
