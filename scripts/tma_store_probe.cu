// tma_store_probe.cu -- experiment: which tensor-map shapes does UTMASTG accept for copying a run of u64 / u32 elements
// from shared memory to an arbitrary element offset of a global array?  One variant per process (an illegal instruction
// kills the context).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -o build/tma_store_probe scripts/tma_store_probe.cu
// Run: build/tma_store_probe VARIANT   (0..N-1; prints OK / the CUDA error)
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

typedef CUresult (*EncodeTiledFn) (CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                   const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct alignas (64) Maps { CUtensorMap m[4]; };

__device__ __forceinline__ uint32_t smem_u32 (const void *p) { return (uint32_t) __cvta_generic_to_shared (p); }

template <int RANK>
__global__ void probe_kernel (const __grid_constant__ Maps maps, const CUtensorMap *gmaps, int use_global, int lanes, uint32_t c0, int elem_bytes, int box)
{
  extern __shared__ __align__ (128) unsigned char smem[];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) reinterpret_cast<uint32_t *> (smem)[i] = 0x1000u + i;
  asm volatile ("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads ();
  if ((int) threadIdx.x < lanes) {
    const int q = threadIdx.x;
    const CUtensorMap *tm = use_global ? &gmaps[q & 3] : &maps.m[q & 3];
    const void *src = smem + (size_t) q * box * elem_bytes;
    const int x = (int) (c0 + q * box);
    if (RANK == 1)
      asm volatile ("cp.async.bulk.tensor.1d.global.shared::cta.bulk_group [%0, {%1}], [%2];" :: "l"(tm), "r"(x), "r"(smem_u32 (src)) : "memory");
    else
      asm volatile ("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" :: "l"(tm), "r"(x), "r"(0), "r"(smem_u32 (src)) : "memory");
    asm volatile ("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile ("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

int main (int argc, char **argv)
{
  const int variant = argc > 1 ? atoi (argv[1]) : 0;
  // variant bits: 0: rank 2 (else 1); 1: u32 elements (else u64); 2: maps in global memory; 3: 4 lanes issue (else 1);
  // 4: box 16 (u64) / 32 (u32) instead of 256; 5: inner dimension = exact array size instead of 2^30
  const int rank = (variant & 1) ? 2 : 1, eb = (variant & 2) ? 4 : 8, use_global = (variant >> 2) & 1, lanes = (variant & 8) ? 4 : 1;
  const int box = (variant & 16) ? (eb == 8 ? 16 : 32) : 256;
  const uint64_t n = 1 << 20;
  void *p = nullptr;
  EncodeTiledFn encode = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaFree (0);
  if (cudaGetDriverEntryPoint ("cuTensorMapEncodeTiled", (void **) &encode, cudaEnableDefault, &q) != cudaSuccess || !encode) { printf ("variant %d: no entry point\n", variant); return 1; }
  unsigned char *d = nullptr;
  cudaMalloc (&d, n * eb);
  cudaMemset (d, 0, n * eb);
  Maps maps;
  const cuuint64_t inner = (variant & 32) ? n : (1ull << 30);
  const cuuint64_t dims[2] = {inner, 1};
  const cuuint64_t strides[1] = {inner * (cuuint64_t) eb};
  const cuuint32_t bx[2] = {(cuuint32_t) box, 1}, es[2] = {1, 1};
  for (int m = 0; m < 4; m++) {
    CUresult r = encode (&maps.m[m], eb == 8 ? CU_TENSOR_MAP_DATA_TYPE_UINT64 : CU_TENSOR_MAP_DATA_TYPE_UINT32, rank, d, dims, strides, bx, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf ("variant %d: encode failed %d\n", variant, (int) r); return 1; }
  }
  CUtensorMap *gm = nullptr;
  cudaMalloc (&gm, sizeof (maps));
  cudaMemcpy (gm, &maps, sizeof (maps), cudaMemcpyHostToDevice);
  const uint32_t c0 = argc > 2 ? (uint32_t) atoi (argv[2]) : 1001;     // default: an odd element offset, no 16-byte alignment
  cudaFuncSetAttribute (probe_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
  cudaFuncSetAttribute (probe_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
  if (rank == 1) probe_kernel<1><<<1, 128, 16384>>> (maps, gm, use_global, lanes, c0, eb, box);
  else probe_kernel<2><<<1, 128, 16384>>> (maps, gm, use_global, lanes, c0, eb, box);
  cudaError_t e = cudaDeviceSynchronize ();
  if (e != cudaSuccess) { printf ("variant %d c0 %u (rank %d, %d-byte elements, box %d, maps in %s, %d lanes, inner %llu): %s\n", variant, c0, rank, eb, box, use_global ? "global" : "param", lanes, (unsigned long long) inner, cudaGetErrorString (e)); return 1; }
  // check
  const size_t words = (size_t) lanes * box * eb / 4;
  uint32_t *h = (uint32_t *) malloc (n * eb);
  cudaMemcpy (h, d, n * eb, cudaMemcpyDeviceToHost);
  size_t bad = 0;
  const size_t w0 = (size_t) c0 * eb / 4;
  for (size_t i = 0; i < n * eb / 4; i++) {
    const uint32_t want = (i >= w0 && i < w0 + words) ? 0x1000u + (uint32_t) (i - w0) : 0u;
    if (h[i] != want) bad++;
  }
  printf ("variant %d c0 %u (rank %d, %d-byte elements, box %d, maps in %s, %d lanes, inner %llu): %s\n", variant, c0, rank, eb, box, use_global ? "global" : "param", lanes,
          (unsigned long long) inner, bad ? "WRONG DATA" : "OK");
  return bad != 0;
}
