// oracle/emd_ref_harness.cu -- TEST INFRASTRUCTURE ONLY (builds into the git-ignored oracle/_ref/).
//
// Runs the reference's own auction-EMD kernels AND its own driver loop (metrics/emd/emd_cuda.cu:23-282,
// `emd_cuda_forward`) on raw device buffers: the reference source is #included from where it lies
// (-DSPGAN_REF_EMD_SRC="..."), with oracle/aten_shim standing in for ATen.  Work arrays are initialised exactly as
// metrics/emd/emd_module.py:45-58 does (assignment = assignment_inv = -1, everything else 0).
//
//   emd_ref_harness <in.bin> <out.bin>
//   in.bin : int32 B, int32 n, float32 eps, int32 iters, float32 xyz1[B*n*3], float32 xyz2[B*n*3]
//   out.bin: float32 dist[B*n], int32 assignment[B*n]
#include <cstdio>
#include <cstdlib>
#include <vector>

#include SPGAN_REF_EMD_SRC

#define CK(x)                                                                  \
    do {                                                                       \
        cudaError_t e_ = (x);                                                  \
        if (e_ != cudaSuccess) {                                               \
            fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_));           \
            return 2;                                                          \
        }                                                                      \
    } while (0)

template <typename T>
static at::Tensor dev(size_t count, int fill_byte, int64_t d0 = 0, int64_t d1 = 0, int64_t d2 = 0) {
    at::Tensor t;
    cudaMalloc(&t.ptr, count * sizeof(T));
    cudaMemset(t.ptr, fill_byte, count * sizeof(T));
    t.dims[0] = d0; t.dims[1] = d1; t.dims[2] = d2;
    return t;
}

int main(int argc, char** argv) {
    if (argc != 3) { fprintf(stderr, "usage: %s in.bin out.bin\n", argv[0]); return 1; }
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 1; }
    int B, n, iters; float eps;
    if (fread(&B, 4, 1, f) != 1 || fread(&n, 4, 1, f) != 1 || fread(&eps, 4, 1, f) != 1 || fread(&iters, 4, 1, f) != 1) return 1;
    const size_t cnt = (size_t)B * n;
    std::vector<float> h1(cnt * 3), h2(cnt * 3);
    if (fread(h1.data(), 4, cnt * 3, f) != cnt * 3 || fread(h2.data(), 4, cnt * 3, f) != cnt * 3) return 1;
    fclose(f);

    at::Tensor xyz1 = dev<float>(cnt * 3, 0, B, n, 3), xyz2 = dev<float>(cnt * 3, 0, B, n, 3);
    CK(cudaMemcpy(xyz1.ptr, h1.data(), cnt * 12, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(xyz2.ptr, h2.data(), cnt * 12, cudaMemcpyHostToDevice));
    at::Tensor dist = dev<float>(cnt, 0, B, n);
    at::Tensor assignment = dev<int>(cnt, 0xff, B, n);              // -1
    at::Tensor assignment_inv = dev<int>(cnt, 0xff, B, n);          // -1
    at::Tensor price = dev<float>(cnt, 0, B, n);
    at::Tensor bid = dev<int>(cnt, 0, B, n);
    at::Tensor bid_increments = dev<float>(cnt, 0, B, n);
    at::Tensor max_increments = dev<float>(cnt, 0, B, n);
    at::Tensor unass_idx = dev<int>(cnt, 0, (int64_t)cnt);
    at::Tensor max_idx = dev<int>(cnt, 0, (int64_t)cnt);
    at::Tensor unass_cnt = dev<int>(512, 0, 512), unass_cnt_sum = dev<int>(512, 0, 512), cnt_tmp = dev<int>(512, 0, 512);

    const int rc = emd_cuda_forward(xyz1, xyz2, dist, assignment, price, assignment_inv, bid, bid_increments,
                                    max_increments, unass_idx, unass_cnt, unass_cnt_sum, cnt_tmp, max_idx, eps, iters);
    CK(cudaDeviceSynchronize());
    if (rc != 1) { fprintf(stderr, "emd_cuda_forward returned %d\n", rc); return 3; }
    std::vector<float> hd(cnt);
    std::vector<int> ha(cnt);
    CK(cudaMemcpy(hd.data(), dist.ptr, cnt * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(ha.data(), assignment.ptr, cnt * 4, cudaMemcpyDeviceToHost));
    f = fopen(argv[2], "wb");
    if (!f) { perror(argv[2]); return 1; }
    fwrite(hd.data(), 4, cnt, f);
    fwrite(ha.data(), 4, cnt, f);
    fclose(f);
    return 0;
}
