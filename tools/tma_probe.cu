// tools/tma_probe.cu -- development probe: which 3-D TMA box shapes / coordinates does the
// B200 accept for the population tiles?  Usage: tma_probe <elem 4|8> <ly> <nx> <by> <bx> <c0> <c1>
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/tma_probe tools/tma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

typedef CUresult (*EncodeTiled_t)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <typename T>
__global__ void probe(const __grid_constant__ CUtensorMap tmap, int c0, int c1, int nelem, T *out) {
  extern __shared__ __align__(128) unsigned char smem[];
  T *tile = reinterpret_cast<T *>(smem);
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem + (((size_t)nelem * sizeof(T) + 127) / 128) * 128);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"((uint32_t)(nelem * sizeof(T)))
                 : "memory");
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            smem_u32(tile)),
        "l"(&tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(0)
        : "memory");
  }
  __syncthreads();
  uint32_t done;
  do {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(0)
        : "memory");
  } while (!done);
  for (int i = threadIdx.x; i < nelem; i += blockDim.x) out[i] = tile[i];
}

template <typename T>
int run(int ly, int nx, int by, int bx, int c0, int c1) {
  const int pitch = (ly + 31) / 32 * 32;
  const size_t plane = (size_t)nx * pitch;
  std::vector<T> h(plane * 9);
  for (int q = 0; q < 9; ++q)
    for (int x = 0; x < nx; ++x)
      for (int y = 0; y < pitch; ++y) h[q * plane + (size_t)x * pitch + y] = (T)(q * 1000000 + x * 1000 + y);
  T *d, *out;
  cudaMalloc(&d, sizeof(T) * h.size());
  cudaMemcpy(d, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice);
  const int nelem = by * bx * 9;
  cudaMalloc(&out, sizeof(T) * nelem);
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  CUtensorMap tm;
  const cuuint64_t dims[3] = {(cuuint64_t)ly, (cuuint64_t)nx, 9};
  const cuuint64_t strides[2] = {(cuuint64_t)pitch * sizeof(T), (cuuint64_t)plane * sizeof(T)};
  const cuuint32_t box[3] = {(cuuint32_t)by, (cuuint32_t)bx, 9};
  const cuuint32_t es[3] = {1, 1, 1};
  CUresult r = ((EncodeTiled_t)fn)(&tm, sizeof(T) == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d,
                                   dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 2; }
  const size_t smem = (((size_t)nelem * sizeof(T) + 127) / 128) * 128 + 16;
  cudaFuncSetAttribute(probe<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe<T><<<1, 256, smem>>>(tm, c0, c1, nelem, out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 3; }
  std::vector<T> res(nelem);
  cudaMemcpy(res.data(), out, sizeof(T) * nelem, cudaMemcpyDeviceToHost);
  long bad = 0;
  for (int qq = 0; qq < 9; ++qq)
    for (int x = 0; x < bx; ++x)
      for (int y = 0; y < by; ++y) {
        const int gx = c1 + x, gy = c0 + y;
        const T want = (gx >= 0 && gx < nx && gy >= 0 && gy < ly) ? (T)(qq * 1000000 + gx * 1000 + gy) : (T)0;
        if (res[((size_t)qq * bx + x) * by + y] != want) ++bad;
      }
  printf("ok, %ld mismatches of %d\n", bad, nelem);
  return bad ? 4 : 0;
}

int main(int argc, char **argv) {
  if (argc < 8) return 1;
  const int el = atoi(argv[1]), ly = atoi(argv[2]), nx = atoi(argv[3]), by = atoi(argv[4]), bx = atoi(argv[5]),
            c0 = atoi(argv[6]), c1 = atoi(argv[7]);
  printf("elem %d ly %d nx %d box %dx%dx9 at (%d,%d): ", el, ly, nx, by, bx, c0, c1);
  return el == 8 ? run<double>(ly, nx, by, bx, c0, c1) : run<float>(ly, nx, by, bx, c0, c1);
}
