// Family 2, dense strided permute through the Tensor Memory Accelerator.
//
// When the destination's innermost dimension is also contiguous in the source (the permutation only
// reorders outer dimensions: the T1 relayout of BASELINE config 2, every `Permute` that keeps the last
// bond), the whole copy is expressed as two tensor maps -- the SOURCE seen in destination dimension
// order (its strides carry the permutation) and the contiguous DESTINATION -- and a persistent kernel in
// which one thread per CTA drives a 6-deep ring of 32 KB shared-memory tiles:
//     cp.async.bulk.tensor (global -> shared, mbarrier complete_tx)   ... UTMALDG
//     cp.async.bulk.tensor (shared -> global, bulk_group)             ... UTMASTG
// No register or LSU traffic on the data path, edges clipped by the hardware.  Tensors with more than
// five collapsed dimensions, odd strides or an odd inner extent fall back to the LSU kernels in
// relayout.cu (the caller does that when this returns T10_ERR_UNSUPPORTED).
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "t10_common.cuh"

namespace t10 {

constexpr int TMA_STAGES = 6;
constexpr int TMA_TILE_BYTES = 32 * 1024;

struct TmaTiling {
  unsigned box[5];       // innermost first
  unsigned ntile[5];     // tiles per dimension
  unsigned long long ntiles;
  unsigned tile_bytes;
};

__device__ __forceinline__ void tma_mbar_init(uint64_t* bar, unsigned count) {
  unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.init.shared.b64 [%0], %1;\n" ::"r"(a), "r"(count));
}
__device__ __forceinline__ void tma_mbar_expect(uint64_t* bar, unsigned bytes) {
  unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.arrive.expect_tx.shared.b64 _, [%0], %1;\n" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_mbar_wait(uint64_t* bar, unsigned parity) {
  unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "TMA_WAIT:\n"
      "mbarrier.try_wait.parity.shared.b64 p, [%0], %1;\n"
      "@p bra TMA_DONE;\n"
      "bra TMA_WAIT;\n"
      "TMA_DONE:\n"
      "}\n" ::"r"(a), "r"(parity)
      : "memory");
}

__global__ void __launch_bounds__(32, 1)
k_permute_tma(const __grid_constant__ CUtensorMap src_map, const __grid_constant__ CUtensorMap dst_map,
              const __grid_constant__ TmaTiling tl) {
  extern __shared__ __align__(128) unsigned char tma_smem[];
  __shared__ __align__(8) uint64_t full[TMA_STAGES];
  if (threadIdx.x != 0) return;
  for (int s = 0; s < TMA_STAGES; ++s) tma_mbar_init(&full[s], 1);
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");

  const unsigned long long first = blockIdx.x, step = gridDim.x;
  const unsigned long long mine = first < tl.ntiles ? (tl.ntiles - first + step - 1) / step : 0;
  auto coords = [&](unsigned long long i, int* c) {
    unsigned long long t = first + i * step;
#pragma unroll
    for (int d = 0; d < 5; ++d) {
      c[d] = (int)((t % tl.ntile[d]) * tl.box[d]);
      t /= tl.ntile[d];
    }
  };
  auto load = [&](unsigned long long i) {
    const int b = (int)(i % TMA_STAGES);
    int c[5];
    coords(i, c);
    unsigned dst = (unsigned)__cvta_generic_to_shared(tma_smem + (size_t)b * TMA_TILE_BYTES);
    unsigned bar = (unsigned)__cvta_generic_to_shared(&full[b]);
    tma_mbar_expect(&full[b], tl.tile_bytes);
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];\n" ::
            "r"(dst), "l"(&src_map), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]), "r"(c[4]), "r"(bar)
        : "memory");
  };
  for (unsigned long long i = 0; i < mine && i < TMA_STAGES; ++i) load(i);
  for (unsigned long long i = 0; i < mine; ++i) {
    const int b = (int)(i % TMA_STAGES);
    tma_mbar_wait(&full[b], (unsigned)((i / TMA_STAGES) & 1));
    int c[5];
    coords(i, c);
    unsigned srcs = (unsigned)__cvta_generic_to_shared(tma_smem + (size_t)b * TMA_TILE_BYTES);
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    asm volatile(
        "cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];\n" ::"l"(&dst_map),
        "r"(srcs), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]), "r"(c[4])
        : "memory");
    asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
    if (i + TMA_STAGES < mine) {
      asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");  // tile b has left shared memory
      load(i + TMA_STAGES);
    }
  }
  asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// shape / sstride: collapsed dims in DESTINATION order (last = innermost), sstride[last] == 1.
int permute_dense_tma(int dtype, const void* src, void* dst, int rank, const int64_t* shape,
                      const int64_t* sstride, cudaStream_t st) {
  const char* env = getenv("TOR10_PERMUTE_TMA");
  if (env && env[0] == '0') return T10_ERR_UNSUPPORTED;
  const int esz = dtype == T10_F64 ? 8 : 4;
  const int vec = 16 / esz;
  if (rank < 1 || rank > 5) return T10_ERR_UNSUPPORTED;
  if (sstride[rank - 1] != 1) return T10_ERR_UNSUPPORTED;
  if (((uintptr_t)src & 15) || ((uintptr_t)dst & 15)) return T10_ERR_UNSUPPORTED;
  if (shape[rank - 1] % vec) return T10_ERR_UNSUPPORTED;
  int64_t total = 1;
  for (int d = 0; d < rank; ++d) {
    if (shape[d] < 1 || shape[d] >= (1ll << 31)) return T10_ERR_UNSUPPORTED;
    if (d < rank - 1 && (sstride[d] % vec || sstride[d] < 1 || sstride[d] * esz >= (1ll << 40)))
      return T10_ERR_UNSUPPORTED;
    total *= shape[d];
  }
  const char* envmin = getenv("TOR10_PERMUTE_TMA_MIN_BYTES");
  const int64_t min_bytes = envmin ? atoll(envmin) : (1ll << 20);
  if (total * esz < min_bytes) return T10_ERR_UNSUPPORTED;  // tiny tensors: descriptor setup is not worth it
  EncodeTiledFn enc = encode_fn();
  if (!enc) return T10_ERR_UNSUPPORTED;

  // TMA dimension k (innermost first) = destination dimension rank-1-k; pad to 5-D with extent 1
  cuuint64_t gdim[5], sstr[4], dstr[4];
  cuuint32_t box[5], estr[5];
  TmaTiling tl;
  memset(&tl, 0, sizeof(tl));
  int64_t dprod = 1;
  for (int k = 0; k < 5; ++k) {
    const int d = rank - 1 - k;
    gdim[k] = d >= 0 ? (cuuint64_t)shape[d] : 1;
    estr[k] = 1;
    box[k] = 1;
    if (k >= 1) {
      sstr[k - 1] = d >= 0 ? (cuuint64_t)sstride[d] * esz : (cuuint64_t)16;
      dstr[k - 1] = (cuuint64_t)dprod * esz;   // contiguous destination
    }
    dprod *= (int64_t)gdim[k];
  }
  // padded dims need a legal (multiple of 16, non-zero) stride; make them monotone for the encoder
  for (int k = 1; k < 5; ++k) {
    const int d = rank - 1 - k;
    if (d < 0) sstr[k - 1] = dstr[k - 1];
  }
  const int tile_elems = TMA_TILE_BYTES / esz;
  // fill the 32 KB tile budget from the innermost dimension outwards (box extents <= 256)
  box[0] = (cuuint32_t)std::min<int64_t>((int64_t)gdim[0], 256);
  box[0] -= box[0] % vec;
  int64_t filled = box[0];
  for (int k = 1; k < 5; ++k) {
    box[k] = (cuuint32_t)std::min<int64_t>(std::min<int64_t>((int64_t)gdim[k], 256), std::max<int64_t>(1, tile_elems / filled));
    filled *= box[k];
  }
  tl.ntiles = 1;
  tl.tile_bytes = esz;
  for (int k = 0; k < 5; ++k) {
    tl.box[k] = box[k];
    tl.ntile[k] = (unsigned)(((int64_t)gdim[k] + box[k] - 1) / box[k]);
    tl.ntiles *= tl.ntile[k];
    tl.tile_bytes *= box[k];
  }
  const CUtensorMapDataType dt = dtype == T10_F64 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUtensorMap smap, dmap;
  CUresult r1 = enc(&smap, dt, 5, const_cast<void*>(src), gdim, sstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CUresult r2 = enc(&dmap, dt, 5, dst, gdim, dstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r1 != CUDA_SUCCESS || r2 != CUDA_SUCCESS) return T10_ERR_UNSUPPORTED;  // e.g. overlapping-stride views

  static bool attr_by_dev[T10_MAX_DEVICES] = {false};
  bool& attr_set = attr_by_dev[cur_device()];
  const int smem = TMA_STAGES * TMA_TILE_BYTES;
  if (!attr_set) {
    T10_CUDA(cudaFuncSetAttribute(k_permute_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  unsigned grid = (unsigned)std::min<unsigned long long>(tl.ntiles, (unsigned long long)sm_count());
  k_permute_tma<<<grid, 32, smem, st>>>(smap, dmap, tl);
  T10_LAUNCH_CHECK();
  return T10_OK;
}

}  // namespace t10
