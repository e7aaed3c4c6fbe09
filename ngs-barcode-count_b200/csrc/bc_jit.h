// bc_jit.h — run-time specialisation of the decode kernel (bc_jit.cu): NVRTC compiles csrc/bc_decode.cuh for sm_100a with
// the run constants (scheme, barcode slots, quality runs, batch geometry) as a constexpr object.
#pragma once
#include <string>

#include "bc_device.cuh"

namespace bc {

struct JitDecode {
    const void* kernel = nullptr;  // cudaKernel_t of k_decode_jit, usable with cudaLaunchKernel on any device
    uint32_t W = 0, plane_stride = 0, qual_stride = 0;  // the batch geometry it was compiled for
};

// The specialised kernel for (cfg, geometry), compiled once per process and configuration.  kernel == nullptr when it is
// not available (no libnvrtc at run time, or a compile error): *why then says why and the caller keeps the generic kernel.
JitDecode jit_decode(const DevCfg& cfg, uint32_t W, uint32_t plane_stride, uint32_t qual_stride, std::string* why);

// compile only (no device needed): size of the cubin, 0 on failure with the reason in *log
size_t jit_compile_check(const DevCfg& cfg, uint32_t W, uint32_t plane_stride, uint32_t qual_stride, std::string* log);

}  // namespace bc
