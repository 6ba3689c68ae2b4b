// oracle/aten_shim/ATen/ATen.h -- TEST INFRASTRUCTURE ONLY.
// A stand-in for <ATen/ATen.h> with just what metrics/emd/emd_cuda.cu touches (Tensor::size, Tensor::data<T>), so
// that the reference's CUDA source compiles from where it lies, unmodified, into a stand-alone harness
// (oracle/emd_ref_harness.cu) without libtorch.  Nothing of the reference is copied: its file is #included.
#pragma once
#include <stdint.h>

namespace at {
struct Tensor {
    void* ptr;
    int64_t dims[3];
    int64_t size(int i) const { return dims[i]; }
    template <typename T>
    T* data() const { return reinterpret_cast<T*>(ptr); }
};
}  // namespace at
