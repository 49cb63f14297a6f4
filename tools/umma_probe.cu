// Hardware probe for the UMMA shared-memory descriptor and TMA layout semantics
// the shift-GEMM kernels rely on.  Host code (tools/umma_probe.py) builds raw
// shared-memory images; this kernel copies an image to smem, issues a chain of
// tcgen05.mma with caller-supplied descriptors, and returns the accumulator.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include "../cmfpy_b200/csrc/sm100_ptx.cuh"

using namespace cmf::ptx;

__global__ void __launch_bounds__(128)
umma_probe_kernel(const uint8_t* __restrict__ image, int nbytes, uint64_t adesc0, uint64_t bdesc0,
                  uint32_t a_adv, uint32_t b_adv, int nsteps, uint32_t idesc, int N, int fix_base_offset,
                  int repeat, float* __restrict__ out, int* __restrict__ status) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid * 16; i < nbytes; i += 128 * 16) *(uint4*)(smem + i) = *(const uint4*)(image + i);
  fence_proxy_async();
  if (warp == 0) { tmem_alloc(&tmem_base_s, 256); tmem_relinquish(); }
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t base = smem_u32(smem);
  if (tid == 0) {
    status[1] = (int)base;
    long long t0 = clock64();
    for (int rep = 0; rep < repeat; ++rep)
    for (int s = 0; s < nsteps; ++s) {
      uint64_t ad = adesc0 + (uint64_t)((base >> 4) + s * a_adv);
      uint64_t bd = bdesc0 + (uint64_t)((base >> 4) + s * b_adv);
      if (fix_base_offset) {   // base_offset = (start_address >> 7) & 7 for both operands
        uint32_t sa = (uint32_t)(ad & 0x3FFF) << 4, sb = (uint32_t)(bd & 0x3FFF) << 4;
        if (fix_base_offset & 1) ad = (ad & ~((uint64_t)7 << 49)) | ((uint64_t)((sa >> 7) & 7) << 49);
        if (fix_base_offset & 2) bd = (bd & ~((uint64_t)7 << 49)) | ((uint64_t)((sb >> 7) & 7) << 49);
      }
      mma_tf32_ss(tmem_base, ad, bd, idesc, (s > 0 || rep > 0) ? 1u : 0u);
    }
    mma_commit(&bar);
    mbar_wait(&bar, 0, 1u << 24);
    status[2] = (int)(clock64() - t0);
  }
  const bool ok = mbar_wait(&bar, 0, 1u << 24);
  tc_fence_after();
  if (ok) {
    for (int c = 0; c < N; c += 32) {
      uint32_t r[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + c, r);
      tmem_ld_wait();
      for (int j = 0; j < 32; ++j) out[(size_t)(warp * 32 + lane) * N + c + j] = __uint_as_float(r[j]);
    }
  }
  if (tid == 0) status[0] = ok ? 1 : -1;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 256);
}

__global__ void __launch_bounds__(128)
tma_probe_kernel(const __grid_constant__ CUtensorMap tmap, int c0, int c1, int box_bytes, int dst_off,
                 uint8_t* __restrict__ dump, int dump_bytes, int* __restrict__ status) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar;
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x;
  for (int i = tid * 16; i < dump_bytes; i += 128 * 16) *(uint4*)(smem + i) = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  fence_proxy_async();
  __syncthreads();
  if (tid == 0) {
    mbar_arrive_expect_tx(&bar, (uint32_t)box_bytes);
    tma_load_2d(smem + dst_off, &tmap, &bar, c0, c1);
  }
  const bool ok = mbar_wait(&bar, 0, 1u << 22);
  __syncthreads();
  for (int i = tid * 16; i < dump_bytes; i += 128 * 16) *(uint4*)(dump + i) = *(const uint4*)(smem + i);
  if (tid == 0) status[0] = ok ? 1 : -1;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

extern "C" {

int probe_umma(const void* image_host, int nbytes, unsigned long long adesc0, unsigned long long bdesc0,
               unsigned a_adv, unsigned b_adv, int nsteps, unsigned idesc, int N, int fix_base_offset,
               int repeat, float* out_host, int* status_host) {
  uint8_t* dimg; float* dout; int* dst;
  cudaMalloc(&dimg, nbytes); cudaMalloc(&dout, 128 * N * 4); cudaMalloc(&dst, 16);
  cudaMemcpy(dimg, image_host, nbytes, cudaMemcpyHostToDevice);
  cudaMemset(dout, 0xFF, 128 * N * 4); cudaMemset(dst, 0, 16);
  int smem = nbytes + 1024;
  cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  umma_probe_kernel<<<1, 128, smem>>>(dimg, nbytes, adesc0, bdesc0, a_adv, b_adv, nsteps, idesc, N, fix_base_offset, repeat, dout, dst);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("probe_umma: %s\n", cudaGetErrorString(e)); return 1; }
  cudaMemcpy(out_host, dout, 128 * N * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(status_host, dst, 12, cudaMemcpyDeviceToHost);
  cudaFree(dimg); cudaFree(dout); cudaFree(dst);
  return 0;
}

// 2-D fp32 tensor map: global [rows][cols] with row pitch `pitch_bytes`; box = (box_cols, box_rows)
int probe_tma(const float* g_host, int rows, int cols, int box_cols, int box_rows, int swizzle /*0 none,1 32B,2 64B,3 128B*/,
              int c0, int c1, int dst_off, unsigned char* dump_host, int dump_bytes, int* status_host) {
  EncodeTiledFn encode = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres) != cudaSuccess || !encode) {
    printf("probe_tma: no cuTensorMapEncodeTiled\n"); return 1;
  }
  float* dg; uint8_t* ddump; int* dst;
  cudaMalloc(&dg, (size_t)rows * cols * 4); cudaMalloc(&ddump, dump_bytes); cudaMalloc(&dst, 8);
  cudaMemcpy(dg, g_host, (size_t)rows * cols * 4, cudaMemcpyHostToDevice); cudaMemset(dst, 0, 8);
  CUtensorMap tmap;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dg, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      (CUtensorMapSwizzle)swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("probe_tma: encode failed %d\n", (int)r); return 2; }
  int smem = dump_bytes + 1024;
  cudaFuncSetAttribute(tma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  tma_probe_kernel<<<1, 128, smem>>>(tmap, c0, c1, box_cols * box_rows * 4, dst_off, ddump, dump_bytes, dst);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("probe_tma: %s\n", cudaGetErrorString(e)); return 1; }
  cudaMemcpy(dump_host, ddump, dump_bytes, cudaMemcpyDeviceToHost);
  cudaMemcpy(status_host, dst, 8, cudaMemcpyDeviceToHost);
  cudaFree(dg); cudaFree(ddump); cudaFree(dst);
  return 0;
}

}  // extern "C"
