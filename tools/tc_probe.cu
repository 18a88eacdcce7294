// Bring-up probe for tcgen05.mma kind::tf32 with no-swizzle shared-memory descriptors (development tool,
// not part of the library).  The host builds the raw shared-memory images of A and B and the descriptor
// fields; the kernel copies the images to shared memory, issues ONE MMA (M=128, N=128, K=8), commits,
// waits with a bounded spin and dumps the 128x128 fp32 accumulator.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/tc_probe tools/tc_probe.cu && /tmp/tc_probe
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x)                                                                         \
  do {                                                                                \
    cudaError_t e = (x);                                                              \
    if (e != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__);  \
      exit(1);                                                                        \
    }                                                                                 \
  } while (0)

struct Params {
  unsigned a_bytes, b_bytes;     // image sizes (multiples of 16)
  unsigned long long a_hi, b_hi; // descriptor bits 16..63 (everything but the start address)
  unsigned idesc;
  int issue_mode;                // 0: tid 0 in a divergent branch, 1: elect.sync in converged warp 0
  int prefill;                   // 1: fill TMEM with lane*1000+col before the MMA and accumulate
  int fence_mode;                // 0: fence.proxy.async + syncthreads, 1: none
};

__global__ void __launch_bounds__(128, 1)
k_probe(const float* __restrict__ a_img, const float* __restrict__ b_img, float* __restrict__ out, int* status, Params p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ unsigned tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const unsigned smem0 = (unsigned)__cvta_generic_to_shared(smem);
  const unsigned a_addr = smem0, b_addr = smem0 + ((p.a_bytes + 1023u) & ~1023u);
  if (warp == 0) {
    unsigned dst = (unsigned)__cvta_generic_to_shared(&tmem_base_s);
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(dst), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  if (tid == 0) {
    unsigned a = (unsigned)__cvta_generic_to_shared(&bar);
    asm volatile("mbarrier.init.shared.b64 [%0], 1;\n" ::"r"(a) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  for (unsigned i = tid; i < p.a_bytes / 4; i += 128) reinterpret_cast<float*>(smem)[i] = a_img[i];
  for (unsigned i = tid; i < p.b_bytes / 4; i += 128) reinterpret_cast<float*>(smem + (b_addr - smem0))[i] = b_img[i];
  if (p.fence_mode == 0) asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const unsigned tmem_d = tmem_base_s;
  if (p.prefill) {
    for (int cb = 0; cb < 128; cb += 8) {
      const unsigned taddr = tmem_d + ((unsigned)(warp * 32) << 16) + (unsigned)cb;
      unsigned w[8];
      for (int j = 0; j < 8; ++j) w[j] = __float_as_uint((float)((warp * 32 + lane) * 1000 + cb + j));
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"r"(taddr),
                   "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
                   : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  }
  const uint64_t adesc = (uint64_t)((a_addr >> 4) & 0x3fff) | (p.a_hi << 16);
  const uint64_t bdesc = (uint64_t)((b_addr >> 4) & 0x3fff) | (p.b_hi << 16);
  const unsigned accum = p.prefill ? 1u : 0u;
  const unsigned barp = (unsigned)__cvta_generic_to_shared(&bar);
  bool issuer = false;
  if (p.issue_mode == 0) {
    issuer = tid == 0;
  } else if (warp == 0) {
    unsigned pred;
    asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}\n" : "=r"(pred));
    issuer = pred != 0;
  }
  if (issuer) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(p.idesc), "r"(accum), "r"(0u)
        : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(barp) : "memory");
  }
  __syncwarp();
  bool ok = false;
  for (int spin = 0; spin < (1 << 22) && !ok; ++spin) {
    unsigned r;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(r)
        : "r"(barp), "r"(0u)
        : "memory");
    ok = r != 0;
  }
  if (!ok && tid == 0) status[0] = 1;
  if (tid == 0) status[1] = (int)tmem_d;
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  for (int cb = 0; cb < 128; cb += 8) {
    unsigned r[8];
    const unsigned taddr = tmem_d + ((unsigned)(warp * 32) << 16) + (unsigned)cb;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
    for (int j = 0; j < 8; ++j) out[(warp * 32 + lane) * 128 + cb + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_d), "r"(128u) : "memory");
}

// ---- host ----------------------------------------------------------------------------------------------------
static std::vector<float> A(128 * 8), B(8 * 128), Cref(128 * 128);

// K-major operand (rows x 8 k), no swizzle: 16-byte chunk (row r, k-chunk c) at (r/8)*sbo + c*lbo + (r%8)*16
static std::vector<float> image_kmajor(const float* X /* [rows][8] */, int rows, unsigned lbo, unsigned sbo, unsigned* bytes) {
  unsigned total = (rows / 8 - 1) * sbo + lbo + 128;
  std::vector<float> img((total + 15) / 16 * 4 + 64, 0.f);
  for (int r = 0; r < rows; ++r)
    for (int c = 0; c < 2; ++c) {
      unsigned off = (r / 8) * sbo + c * lbo + (r % 8) * 16;
      for (int j = 0; j < 4; ++j) img[off / 4 + j] = X[r * 8 + c * 4 + j];
    }
  *bytes = (unsigned)img.size() * 4;
  return img;
}
// MN-major operand (8 k x cols), no swizzle: chunk (k row, 4-col chunk c) at c*sbo + (k%8)*16   (one k group)
static std::vector<float> image_mnmajor(const float* X /* [8][cols] */, int cols, unsigned sbo, unsigned* bytes) {
  unsigned total = (cols / 4 - 1) * sbo + 128;
  std::vector<float> img((total + 15) / 16 * 4 + 64, 0.f);
  for (int k = 0; k < 8; ++k)
    for (int c = 0; c < cols / 4; ++c) {
      unsigned off = c * sbo + k * 16;
      for (int j = 0; j < 4; ++j) img[off / 4 + j] = X[k * cols + c * 4 + j];
    }
  *bytes = (unsigned)img.size() * 4;
  return img;
}

static unsigned long long hi_bits(unsigned lbo, unsigned sbo, int version, int layout) {
  unsigned long long d = 0;
  d |= (unsigned long long)((lbo >> 4) & 0x3fff);
  d |= (unsigned long long)((sbo >> 4) & 0x3fff) << 16;
  d |= (unsigned long long)version << 30;
  d |= (unsigned long long)layout << 45;
  return d;   // shifted left by 16 in the kernel
}

static void run(const char* name, const std::vector<float>& aimg, const std::vector<float>& bimg, Params p, bool transposedB = false) {
  float *da, *db, *dout;
  int* dst;
  CK(cudaMalloc(&da, aimg.size() * 4));
  CK(cudaMalloc(&db, bimg.size() * 4));
  CK(cudaMalloc(&dout, 128 * 128 * 4));
  CK(cudaMalloc(&dst, 8));
  CK(cudaMemcpy(da, aimg.data(), aimg.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, bimg.data(), bimg.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dout, 0xff, 128 * 128 * 4));
  CK(cudaMemset(dst, 0, 8));
  p.a_bytes = (unsigned)aimg.size() * 4;
  p.b_bytes = (unsigned)bimg.size() * 4;
  int smem = ((p.a_bytes + 1023) & ~1023) + p.b_bytes + 1024;
  CK(cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  k_probe<<<1, 128, smem>>>(da, db, dout, dst, p);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("%-40s CUDA error: %s\n", name, cudaGetErrorString(e));
    exit(2);
  }
  std::vector<float> out(128 * 128);
  int st[2];
  CK(cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(st, dst, 8, cudaMemcpyDeviceToHost));
  double maxerr = 0, maxabs = 0;
  int nz = 0;
  for (int i = 0; i < 128 * 128; ++i) {
    double ref = Cref[i] + (p.prefill ? (double)((i / 128) * 1000 + (i % 128)) : 0.0);
    double d = out[i] - ref;
    if (d < 0) d = -d;
    if (d > maxerr) maxerr = d;
    if (out[i] != 0.f) ++nz;
    double a = out[i] < 0 ? -out[i] : out[i];
    if (a > maxabs) maxabs = a;
  }
  printf("%-44s timeout=%d tmem=0x%x nonzero=%5d maxabs=%9.3f maxerr=%9.4f  C[0,0..3]=%g %g %g %g  C[9,5]=%g (ref %g)\n", name,
         st[0], st[1], nz, maxabs, maxerr, out[0], out[1], out[2], out[3], out[9 * 128 + 5], Cref[9 * 128 + 5]);
  cudaFree(da);
  cudaFree(db);
  cudaFree(dout);
  cudaFree(dst);
}

int main() {
  srand(1);
  for (auto& x : A) x = (float)((rand() % 17) - 8) * 0.25f;   // exactly representable in tf32
  for (auto& x : B) x = (float)((rand() % 13) - 6) * 0.5f;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < 128; ++n) {
      float s = 0;
      for (int k = 0; k < 8; ++k) s += A[m * 8 + k] * B[k * 128 + n];
      Cref[m * 128 + n] = s;
    }
  std::vector<float> Bt(128 * 8);   // B transposed: [n][k], K-major
  for (int k = 0; k < 8; ++k)
    for (int n = 0; n < 128; ++n) Bt[n * 8 + k] = B[k * 128 + n];

  const unsigned idesc_base = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
  unsigned ab, bb;
  // --- variant family 1: both operands K-major (the best documented form)
  for (int issue = 0; issue < 2; ++issue)
    for (int version = 1; version >= 0; --version) {
      // dense packing: core matrix = 128 B; k-chunks adjacent (LBO=128), 8-row groups at SBO=256
      auto ai = image_kmajor(A.data(), 128, 128, 256, &ab);
      auto bi = image_kmajor(Bt.data(), 128, 128, 256, &bb);
      Params p{};
      p.a_hi = hi_bits(128, 256, version, 0);
      p.b_hi = hi_bits(128, 256, version, 0);
      p.idesc = idesc_base;
      p.issue_mode = issue;
      char nm[96];
      snprintf(nm, sizeof nm, "KK lbo=128 sbo=256 ver=%d issue=%d", version, issue);
      run(nm, ai, bi, p);
      // swapped roles of LBO / SBO in the descriptor (image unchanged)
      p.a_hi = hi_bits(256, 128, version, 0);
      p.b_hi = hi_bits(256, 128, version, 0);
      snprintf(nm, sizeof nm, "KK desc swapped lbo/sbo ver=%d issue=%d", version, issue);
      run(nm, ai, bi, p);
    }
  // prefill + accumulate: does the MMA touch TMEM at all?
  {
    auto ai = image_kmajor(A.data(), 128, 128, 256, &ab);
    auto bi = image_kmajor(Bt.data(), 128, 128, 256, &bb);
    Params p{};
    p.a_hi = hi_bits(128, 256, 1, 0);
    p.b_hi = hi_bits(128, 256, 1, 0);
    p.idesc = idesc_base;
    p.prefill = 1;
    run("KK prefill+accumulate", ai, bi, p);
    p.prefill = 0;
    p.fence_mode = 1;
    run("KK no proxy fence", ai, bi, p);
  }
  // the layout the library kernel uses for A: LBO=128, SBO=1024 (BK=32 stage), still K=8 per MMA
  {
    auto ai = image_kmajor(A.data(), 128, 128, 1024, &ab);
    auto bi = image_kmajor(Bt.data(), 128, 128, 1024, &bb);
    Params p{};
    p.a_hi = hi_bits(128, 1024, 1, 0);
    p.b_hi = hi_bits(128, 1024, 1, 0);
    p.idesc = idesc_base;
    run("KK lbo=128 sbo=1024", ai, bi, p);
  }
  // padded (bank-conflict-free) strides the library kernel wants: A lbo=144 sbo=1152, B lbo=128 sbo=1040
  {
    auto ai = image_kmajor(A.data(), 128, 144, 1152, &ab);
    auto bi = image_kmajor(Bt.data(), 128, 128, 1040, &bb);
    Params p{};
    p.a_hi = hi_bits(144, 1152, 1, 0);
    p.b_hi = hi_bits(128, 1040, 1, 0);
    p.idesc = idesc_base;
    run("KK padded A(144,1152) B(128,1040)", ai, bi, p);
  }
  // --- variant family 2: B MN-major as in the library kernel (SBO=128 between 4-col chunks, LBO = next k group)
  for (int swap = 0; swap < 2; ++swap) {
    auto ai = image_kmajor(A.data(), 128, 128, 1024, &ab);
    auto bi = image_mnmajor(B.data(), 128, 128, &bb);
    Params p{};
    p.a_hi = hi_bits(128, 1024, 1, 0);
    p.b_hi = swap ? hi_bits(128, 4096, 1, 0) : hi_bits(4096, 128, 1, 0);
    p.idesc = idesc_base | (1u << 16);
    run(swap ? "A K-major, B MN-major (lbo/sbo swapped)" : "A K-major, B MN-major (library layout)", ai, bi, p);
  }
  return 0;
}
