/*
 * tor10_b200.h -- C-ABI of the B200 (sm_100a) kernels behind tor10's block-sparse
 * contraction path.  Built as tor10_b200/lib/libtor10_b200.so (nvcc, no torch, no
 * Python in the signatures: plain pointers, sizes and an opaque stream handle).
 *
 * Each entry point replaces one interpreted-Python / torch call site of the reference
 * (kaihsin/Tor10, paths relative to /root/reference):
 *
 *   family 1 (integer, bit-exact)
 *     t10_fuse_qnums        Bond.combine / UniTensor.GetTotalQnums   Bond.py:310-421, UniTensor.py:2859-2868
 *     t10_blockmap_build    UniTensor.__init__ symmetric branch      UniTensor.py:393-441, Bond.py:13-28, :423-447
 *     t10_unravel           _fx_decompress_idx                        UniTensor.py:24-29
 *     t10_relayout_tables   index algebra of Contiguous()             UniTensor.py:2108-2120
 *   family 2 (HBM-bound data movement)
 *     t10_relayout_blocks   element loop of Contiguous()              UniTensor.py:2108-2129
 *                           (gather form = GetBlock non-contiguous    UniTensor.py:3255-3273)
 *     t10_permute_dense     Tensor.contiguous() / tensordot's copy    UniTensor.py:2057, :2144, :3769
 *     t10_diag_scale        torch.diag + tensordot with a diagonal    UniTensor.py:3756-3765
 *   family 3 (FP64 DMMA / FP32 grouped GEMM)
 *     t10_ggemm             per-sector torch.matmul loop              UniTensor.py:3727-3741
 *                           dense tensordot / Matmul GEMM             UniTensor.py:3769, linalg.py:722-733
 *
 * Conventions
 *   - every function returns T10_OK (0) or an error code; t10_last_error() gives the text
 *     (thread-local).  No exceptions cross the ABI.
 *   - pointers prefixed d_ are DEVICE pointers, h_ are HOST pointers.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *   - the library never owns user data; scratch needed while BUILDING maps is
 *     stream-ordered (cudaMallocAsync) and freed before return.  Hot-path calls
 *     (relayout / permute / ggemm) allocate nothing and do not synchronise.
 *   - bond tags: +1 = BRA, -1 = KET (Bond.py:41).
 */
#ifndef TOR10_B200_H
#define TOR10_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define T10_ABI_VERSION 1
#define T10_MAX_RANK 16          /* bonds per tensor                                   */
#define T10_MAX_NSYM 8           /* symmetries per bond                                */
#define T10_KEYSPACE_MAX (1ll << 26) /* product of per-symmetry total-charge ranges    */
#define T10_ARENA_ALIGN 16       /* sector base alignment in ELEMENTS (128 B for f64)  */

enum t10_status {
  T10_OK = 0,
  T10_ERR_INVALID = 1,     /* bad argument                                             */
  T10_ERR_CUDA = 2,        /* a CUDA runtime call failed                               */
  T10_ERR_KEYSPACE = 3,    /* total-charge key space exceeds T10_KEYSPACE_MAX          */
  T10_ERR_NOBLOCK = 4,     /* row and column spaces share no sector (UniTensor.py:415) */
  T10_ERR_UNSUPPORTED = 5  /* dtype / size outside what the kernels cover              */
};

enum t10_dtype { T10_F64 = 0, T10_F32 = 1 };
enum t10_symkind { T10_SYM_U1 = 0, T10_SYM_ZN = 1 };

typedef void* t10_stream_t;

int t10_version(void);
const char* t10_last_error(void);
/* sm count, compute capability and opt-in shared memory of `device` */
int t10_device_props(int device, int* sm_count, int* cc_major, int* cc_minor, int* smem_optin);

/* ------------------------------------------------------------------ family 1 ---- */

/* Total-charge table of a bond space: out[x, s] for x = row-major flat index over
 * `nbonds` bonds.  Charges of bond b are first multiplied by signs[b] (+1/-1, no modular
 * reduction: Bond.py:469-477), then folded left to right with U1: a+b, Zn: (a+b) mod n
 * with the non-negative numpy remainder (Symmetry.py:33-34, :83-84).  A single bond is not
 * passed through the rule.  h_qoff[b] = row offset of bond b inside d_qnums.            */
int t10_fuse_qnums(int nbonds, int nsym, const int64_t* h_dims, const int32_t* h_signs,
                   const int32_t* h_sym_kind, const int64_t* h_sym_n, const int64_t* h_qoff,
                   const int64_t* d_qnums, int64_t* d_out, t10_stream_t stream);

/* The block map of a symmetric tensor (bonds in MEMORY order, `rowrank` of them in row
 * space).  R = prod dims[:rowrank], C = prod dims[rowrank:].
 *   d_block_qnums [cap_sec, nsym]  sector charges, ascending lexicographic
 *   d_ket_map     [R, 2]           (sector, offset) of every flat row, (-1,-1) if none
 *   d_bra_map     [C, 2]           same for columns
 *   d_ket_idx     [R]              flat rows grouped by sector, ascending inside a sector
 *   d_bra_idx     [C]              same for columns
 *   d_sec_info    [cap_sec, 8]     {rows, cols, ket_start, bra_start, arena_off, ld, 0,0}; ld = cols rounded up
 *                                  to even = leading dimension of the stored sector (16-byte aligned fp64 rows)
 *   d_rowbase     [R]              arena element offset of the row's first element, -1 if none
 *   d_coloff      [C]              column offset inside the sector, -1 if none
 *   h_counts      [4]              {nsec, arena_elems (incl. alignment padding), n_ket_phys, n_bra_phys}
 * cap_sec must be >= min(R, C).  Synchronises `stream` before returning.               */
int t10_blockmap_build(int rank, int rowrank, int nsym, const int64_t* h_dims,
                       const int32_t* h_braket, const int32_t* h_sym_kind,
                       const int64_t* h_sym_n, const int64_t* h_qoff, const int64_t* d_qnums,
                       int64_t cap_sec, int64_t* d_block_qnums, int64_t* d_ket_map,
                       int64_t* d_bra_map, int64_t* d_ket_idx, int64_t* d_bra_idx,
                       int64_t* d_sec_info, int64_t* d_rowbase, int32_t* d_coloff,
                       int64_t* h_counts, t10_stream_t stream);

/* out[t, p] = (idx[t] / strides[p]) mod-chain: multi-index of flat idx[t] for the given
 * row-major strides (UniTensor.py:24-29).                                              */
int t10_unravel(int64_t n, int nstrides, const int64_t* h_strides, const int64_t* d_idx,
                int64_t* d_out, t10_stream_t stream);

/* Relayout address tables for destination layout L read from source layout M.
 * For t in [0,n): x = d_idx[t] is a flat index over `nb` bonds with dims h_dims; every
 * digit i_p contributes i_p*h_srow[p] to the source flat row and i_p*h_scol[p] to the
 * source flat column.  Writes d_to_row[t], d_to_col[t].                                */
int t10_relayout_tables(int64_t n, int nb, const int64_t* h_dims, const int64_t* h_srow,
                        const int64_t* h_scol, const int64_t* d_idx, int32_t* d_to_row,
                        int32_t* d_to_col, t10_stream_t stream);

/* ------------------------------------------------------------------ family 2 ---- */

/* Block-sparse strided permute (gather form).  For every destination sector s, row r,
 * column c:   dst[off_s + r*ld_s + c] = src[rowbase[RR[ks+r]+CR[bs+c]] + coloff[RC[ks+r]+CC[bs+c]]]
 * (0 where the source element is not stored).  The destination is cut into tiles of at most
 * T10_RELAYOUT_TR rows x T10_RELAYOUT_TC columns; a tile record is self-contained:
 *   dst   = off_s + r0*ld_s + c0   element offset of the tile's first element in d_dst
 *   rbase = ks + r0, cbase = bs + c0   first entries of the tile in RR/RC and CR/CC
 *   nr, nc = rows / columns of the tile, dld = ld_s
 * d_src_ket_map / d_src_bra_map (the source layout's [R,2] / [C,2] maps) may be NULL; when
 * given, an element is copied only if its source row and column lie in the SAME sector
 * (needed for Zn layouts whose sign-flipped single-bond spaces lose sectors, quirk Q3).   */
#define T10_RELAYOUT_TR 32
#define T10_RELAYOUT_TC 128
typedef struct t10_relayout_tile {
  int64_t dst;
  int32_t rbase, cbase, nr, nc, dld, pad;
} t10_relayout_tile;
int t10_relayout_blocks(int dtype, const void* d_src, void* d_dst, int64_t ntiles,
                        const t10_relayout_tile* d_tiles, const int32_t* d_RR, const int32_t* d_RC,
                        const int32_t* d_CR, const int32_t* d_CC, const int64_t* d_rowbase,
                        const int32_t* d_coloff, const int64_t* d_src_ket_map,
                        const int64_t* d_src_bra_map, int32_t* d_emit_index, t10_stream_t stream);

/* Replay of a relayout through its element index.  t10_relayout_blocks with d_emit_index != NULL
 * writes, instead of data, the source element index of every destination element it covers into
 * d_emit_index (same addressing as d_dst; -1 where the source is not stored).  t10_relayout_index then
 * performs dst[i] = src[index[i]] for i < n (index -1: dst[i] = 0; -2: dst[i] untouched).  n even,
 * d_dst 16-byte aligned.  Costs 4 B of plan memory and 4 B of extra traffic per element, and cuts
 * the dependent-load chain of the relayout from four loads to two.                           */
int t10_relayout_index(int dtype, const void* d_src, void* d_dst, int64_t n, const int32_t* d_index,
                       t10_stream_t stream);

/* Replay of a relayout through RUNS of its element index (maximal pieces with index[i+1] == index[i]+1, or
 * constant -1 / -2).  d_runs: int32 pairs {dst_start, src_start} sorted by dst_start, nruns real entries
 * followed by the sentinel {n, 0}; d_chunk_first[c] = run containing destination element 1024*c, for
 * c = 0 .. ceil(n/1024) (the last entry = nruns-1).  Same result as t10_relayout_index, without the 4 B per
 * element of index traffic; built by the host from the emitted index (plan time).  d_dst 16-byte aligned. */
int t10_relayout_runs(int dtype, const void* d_src, void* d_dst, int64_t n, const int32_t* d_runs,
                      int64_t nruns, const int32_t* d_chunk_first, t10_stream_t stream);

/* Replay of a relayout through SEGMENTS of its element index: maximal pieces contiguous on both sides, cut at
 * T10_RELAYOUT_SEG_MAX elements.  d_seg: int32 quadruples {dst_start, src_start, len, source} (16-byte aligned);
 * the piece is read from arena h_src[source] (nsrc <= 8 base pointers: this GPU's arena, and in sector-resident
 * multi-GPU runs the NVLink peer mappings of the other ranks' arenas -- the relayout of a sharded tensor fetches
 * what it needs from its owners inside the copy, without a collective).  src_start < 0: zero fill.  One warp per
 * segment; elements no segment covers are left untouched.  Opens the programmatic launch window of the kernel
 * that follows on the stream (t10_ggemm_tma with pdl = 1).                                            */
#define T10_RELAYOUT_SEG_MAX 256
int t10_relayout_segments(int dtype, int nsrc, const void* const* h_src, void* d_dst, int64_t nseg,
                          const int32_t* d_seg, t10_stream_t stream);

/* Multi-GPU sector gather over NVLink peer memory (no reference counterpart: the reference has no
 * multi-device code).  dst[lo_i .. hi_i) = src_i[lo_i .. hi_i) for nseg <= 64 slices, where src_i are
 * peer-mapped (symmetric-memory) staging buffers of the other ranks; bounds multiples of 16 bytes.
 * One launch pulls from all peers at once.                                                 */
int t10_gather_peers(int dtype, void* d_dst, int nseg, const void* const* h_src, const int64_t* h_lo,
                     const int64_t* h_hi, t10_stream_t stream);

/* Dense strided permute: dst (contiguous, shape = h_shape) [i0..i_{r-1}] =
 * src[sum_k i_k * h_src_strides[k]]  (strides in elements).                            */
int t10_permute_dense(int dtype, const void* d_src, void* d_dst, int rank,
                      const int64_t* h_shape, const int64_t* h_src_strides, t10_stream_t stream);

/* Scaling by a diagonal operand without densifying it: a viewed as [outer, d, inner] (contiguous,
 * row-major), out[o, i, j] = a[o, i, j] * diag[i].  (inner = 1: column scaling of an [outer, d]
 * matrix; outer = 1: row scaling of a [d, inner] matrix.)                                   */
int t10_diag_scale(int dtype, const void* d_a, const void* d_diag, void* d_out, int64_t outer,
                   int64_t d, int64_t inner, t10_stream_t stream);

/* ------------------------------------------------------------------ family 3 ---- */

/* One GEMM of a group: C[m,n] (row-major, ldc) = A[m,k] (row-major, lda) * B[k,n]
 * (row-major, ldb); offsets in ELEMENTS from the operand base pointers.                */
typedef struct t10_gemm_problem {
  int64_t a_off, b_off, c_off;
  int32_t m, n, k;
  int32_t lda, ldb, ldc;
} t10_gemm_problem;

/* Tile decomposition used by t10_ggemm for `tile_cfg` (FP64; all BK = 16):
 *   0: 64x64, 4 stages   1: 128x128   2: 128x64   3: 64x64, 3 stages, 3 CTAs/SM (default)
 *   40: 64x64 and 42: 64x128, aligned-only kernel (every lda/ldb/offset even, bases 16-byte aligned)
 *   100: SIMT cross-check kernel (tests).
 * FP32: cfg 50 (default) = tcgen05 / TMEM kernel, 128x128x32 tiles, 3xTF32 split with chunked
 *   round-to-nearest accumulation (fp32 accuracy); cfg 0 = SIMT.
 * t10_ggemm_slots: number of co-resident CTAs of that kernel on the current device (needs CUDA).
 * t10_ggemm_plan_tiles (host only): fills h_tiles [cap,4] = {problem, m0, n0, 0}; call with
 * h_tiles = NULL to size.  With nslots > 0 the tiles are assigned to `nslots` CTA slots by LPT on a
 * DMMA-count cost model and stored slot-major, h_slot_start [nslots+1] delimiting each slot;
 * with nslots = 0 they are listed heaviest first for a grid-strided walk.                  */
int t10_ggemm_tile_shape(int tile_cfg, int* bm, int* bn);
int t10_ggemm_slots(int dtype, int tile_cfg, int* nslots);
int t10_ggemm_plan_tiles(int tile_cfg, int64_t nprob, const t10_gemm_problem* h_prob, int64_t nslots,
                         int64_t cap, int32_t* h_tiles, int32_t* h_slot_start, int64_t* ntiles);

/* Grouped GEMM over the matched sectors: one persistent launch.  Tile schedule:
 *   d_sched and d_slot_start   SM bins (fp64 kernels 0..3): the slot table built with nslots = number of
 *                        SMs holds one LPT bin per SM; the CTAs resident on an SM pull its tiles through
 *                        d_sched[2 + bin] and steal from other bins when theirs is empty.  d_sched is a
 *                        zero-initialised int32[2 + nslots];
 *   d_sched only         dynamic: CTAs pull tiles from the heaviest-first list (built with nslots = 0)
 *                        through d_sched[0]; d_sched is a zero-initialised int32[2];
 *   d_slot_start only    static LPT slots (grid = nslots);
 *   neither              grid-strided walk (grid = min(ntiles, resident CTAs)).
 * d_sched is owned by the caller and private to one stream (the kernel resets it before it exits).   */
int t10_ggemm(int dtype, int tile_cfg, const void* d_A, const void* d_B, void* d_C,
              int64_t nprob, const t10_gemm_problem* d_prob, int64_t ntiles,
              const int32_t* d_tiles, int64_t nslots, const int32_t* d_slot_start,
              int32_t* d_sched, t10_stream_t stream);

/* Stream-K form of the mixed-shape aligned fp64 kernel (tile_cfg 46).  The table lists PIECES: entry t is k-steps
 * [kt0, kt1) (units of 16) of the tile d_tiles[t] (as produced by t10_ggemm_plan_tiles for cfg 46), described by
 * d_pieces[t] = {kt0, kt1, park_slot, nfix | first_fix << 8}.  A piece with park_slot >= 0 does not contain its
 * tile's last k-step: it writes its partial accumulators to workspace slot park_slot (any index other than
 * nslots, normally its CTA; d_ws holds (largest park_slot + 1) * 4352 doubles, d_flags one more int32 than
 * that) and raises d_flags[park_slot] to `epoch`.  The piece that contains the last k-step
 * (park_slot = -1) adds the nfix parked slots first_fix, first_fix+1, ... in that order and stores the tile, so
 * the result does not depend on timing.  d_slot_start[nslots+1] delimits each CTA's pieces; the host lists a
 * CTA's parked piece first and its waiting piece last, and nslots must not exceed the resident CTAs
 * (t10_ggemm_slots).  d_flags: at least int32[nslots + 1], zero-initialised once; `epoch` must differ from
 * launch to launch (a counter); d_flags[nslots] != 0 afterwards reports a parked partial that never arrived. */
int t10_ggemm_streamk(const void* d_A, const void* d_B, void* d_C, int64_t nprob,
                      const t10_gemm_problem* d_prob, int64_t npieces, const int32_t* d_tiles,
                      const int32_t* d_pieces, int64_t nslots, const int32_t* d_slot_start,
                      void* d_ws, int32_t* d_flags, int32_t epoch, t10_stream_t stream);

/* t10_ggemm_streamk with the fused all-gather of t10_ggemm_allgather: every finished tile (the piece that holds
 * its last k-step, after the fix-up) is staged in shared memory -- the staging tile aliases the pipeline stages,
 * so two CTAs stay resident per SM -- and pushed to the npeers arenas by TMA bulk copies; the last CTA raises
 * the peers' flags to `step`.  Arguments as in the two functions it combines.                              */
int t10_ggemm_streamk_allgather(const void* d_A, const void* d_B, int npeers, int self_index,
                                void* const* h_peer_C, void* const* h_peer_flag, uint64_t step, int32_t* d_done_ctr,
                                int64_t nprob, const t10_gemm_problem* d_prob, int64_t npieces,
                                const int32_t* d_tiles, const int32_t* d_pieces, int64_t nslots,
                                const int32_t* d_slot_start, void* d_ws, int32_t* d_flags, int32_t epoch,
                                t10_stream_t stream);

/* TMA-fed form of t10_ggemm_streamk (tile_cfg 47; same tables, same results bit for bit): a producer warp feeds
 * the four DMMA warps of a CTA through `cp.async.bulk.tensor` boxes and full/empty mbarriers, the ring of stages is
 * carried across pieces, sector edges and the K tail are zero-filled by the TMA unit.  Needs one tensor map per
 * sector and operand: t10_ggemm_tma_encode (HOST call, no launch) writes 2 * nprob maps of 128 bytes into h_maps
 * -- map 2i describes A of problem i (k x m, row pitch lda), map 2i + 1 its B (n x k, row pitch ldb) -- for the
 * operand base pointers d_A / d_B; the caller copies them to 64-byte aligned device memory (d_maps) and may reuse
 * them for as long as the operands keep their addresses.  Every lda / ldb / a_off / b_off must be even and the base
 * pointers 16-byte aligned.  pdl != 0 launches with programmatic stream serialisation: the kernel's set-up overlaps
 * the tail of the preceding kernel on the stream (the block relayout that writes A) and its first operand load
 * waits for that kernel's completion (griddepcontrol.wait).
 * t10_ggemm_tma_slots: co-resident CTAs of the kernel (the largest legal nslots).                             */
int t10_ggemm_tma_slots(int* nslots);
int t10_ggemm_tma_encode(int64_t nprob, const t10_gemm_problem* h_prob, const void* d_A, const void* d_B,
                         void* h_maps);
int t10_ggemm_tma(const void* d_maps, void* d_C, int64_t nprob, const t10_gemm_problem* d_prob, int64_t npieces,
                  const int32_t* d_tiles, const int32_t* d_pieces, int64_t nslots, const int32_t* d_slot_start,
                  void* d_ws, int32_t* d_flags, int32_t epoch, int pdl,
                  int nsig, void* const* h_sig_flags, uint64_t sig_value, int32_t* d_done_ctr,
                  int npush, void* const* h_push, t10_stream_t stream);

/* t10_ggemm_tma with npush > 0 (fused GEMM -> all-gather, "push" transport): every finished tile is also stored,
 * straight from the accumulator registers, into h_push[0..npush) -- the result arenas of the OTHER ranks through
 * their NVLink peer mappings, at the same element offsets as in d_C -- so the transfer rides under the remaining
 * pieces' math; the completion signal (nsig > 0 required) then covers the remote stores, and a rank that has
 * acquired every peer's progress word (t10_wait_flags) holds the whole result.                              */

/* Progress words of sector-resident multi-GPU runs (no reference counterpart: the reference is single-device).
 * Every rank keeps ONE uint64 progress word in symmetric memory: the number of resident operations it has
 * completed.  t10_ggemm_tma with nsig > 0: the last CTA of the launch stores `sig_value` into h_sig_flags[0..nsig)
 * (normally just this rank's own word: a local store) with system-scope release semantics (d_done_ctr:
 * zero-initialised int32 arrival counter, reset by the kernel).  t10_signal_flags does the same from a one-warp
 * kernel (after operations that only READ peer memory).
 * t10_wait_flags: a one-warp kernel that acquires *h_flags[i] >= h_need[i], i < n <= 8, before anything launched
 * after it on the stream runs; h_flags[i] is a rank's progress word, a peer's through its NVLink mapping (the READER
 * polls remotely: a producer nobody waits for never touches the link).  Placed before a kernel that reads sectors
 * another rank produced, or that overwrites memory other ranks may still be reading.  timeout_ms > 0 bounds the
 * wait: on expiry *d_err is set to 1 + i and the kernel TRAPS (the CUDA context fails loudly; nothing downstream
 * consumes missing data); timeout_ms = 0 waits for ever, like NCCL.                                        */
int t10_signal_flags(int nsig, void* const* h_sig_flags, uint64_t value, t10_stream_t stream);
int t10_wait_flags(int n, const void* const* h_flags, const uint64_t* h_need, int64_t timeout_ms, int32_t* d_err,
                   t10_stream_t stream);

/* Deterministic second half of a split-K product: d_out[i] = sum_{s < nslices} d_ws[s * n + i], added in slice
 * order.  A GEMM whose output is a handful of tiles but whose K is long (the norms and expectation values of
 * iTEBD.py: [1 x 4096] x [4096 x 1]) is issued as `nslices` problems of one t10_ggemm launch, each writing its
 * partial product to slice s of a workspace, instead of one CTA walking the whole K alone.               */
int t10_reduce_slices(int dtype, const void* d_ws, void* d_out, int64_t n, int nslices, t10_stream_t stream);

/* Fused grouped GEMM -> all-gather (fp64, kernels 0..3; multi-GPU sector sharding, no reference counterpart:
 * the reference is single-device).  Same as t10_ggemm, but every output tile is stored into ALL `npeers`
 * arenas h_peer_C[0..npeers) -- this rank's staging arena and the peers', mapped through NVLink peer memory
 * (torch symmetric memory) -- at the same element offsets, tile by tile, so the transfer overlaps the
 * remaining tiles' math.  With h_peer_flag != NULL the last CTA to finish raises *h_peer_flag[q] (this
 * rank's uint64 slot in rank q's flag array, peer-mapped) to `step` with system-scope release semantics;
 * d_done_ctr is a zero-initialised int32 owned by the caller (reset by the kernel).
 * self_index: which of the arenas lives in THIS GPU's memory (-1: unknown); the aligned kernels (40, 42, 46,
 * stream-K) write that one with plain vector stores and only the others through the copy engine.
 * t10_wait_copy is the consumer: it waits until d_flags[0..nflags) >= step, then copies n elements
 * d_src -> d_dst (staging arena -> result arena).  *d_err is set if a flag does not arrive in ~2 s.      */
int t10_ggemm_allgather(int tile_cfg, const void* d_A, const void* d_B, int npeers, int self_index,
                        void* const* h_peer_C, void* const* h_peer_flag, uint64_t step, int32_t* d_done_ctr,
                        int64_t nprob, const t10_gemm_problem* d_prob, int64_t ntiles,
                        const int32_t* d_tiles, int64_t nslots, const int32_t* d_slot_start,
                        int32_t* d_sched, t10_stream_t stream);
int t10_wait_copy(int dtype, void* d_dst, const void* d_src, int64_t n, const void* d_flags, int nflags,
                  uint64_t step, int32_t* d_err, t10_stream_t stream);

/* ------------------------------------------------------------------ small SVD ---- */

/* Batched SVD of small matrices, one CTA per matrix: one-sided Jacobi in shared memory (replaces cuSOLVER behind
 * `torch.svd(a.Storage, some=True)` of Svd / Svd_truncate, linalg.py:544-678, for matrices that fit one SM -- the
 * 64 x 64 matrix of every iTEBD.py step costs cuSOLVER 1.8 ms on a B200).  h_prob: nine int64 per matrix
 * {m, n, a_off, row_stride, col_stride, u_off, s_off, v_off, w_off} with m >= n, n <= 64 and (m + n') * n' * sizeof <= 200 KB
 * (n' = n rounded up to even; t10_svd_small_limits); element (r, c) of matrix i is d_A[a_off + r * row_stride +
 * c * col_stride] (a wide matrix is passed transposed through its strides).  Writes U [m, n] row-major at
 * d_U + u_off, the singular values, descending, at d_S + s_off and Vh [n, n] row-major at d_Vh + v_off, so that
 * A = U diag(S) Vh.  A zero singular value gets a zero U column; columns below m eps ||A||_F (the numerical null
 * space) are left as they are: their U columns are not orthonormalised.  d_info[i] (may be NULL): bit 0 = matrix i was still
 * rotating after 40 sweeps, bits 1.. = sweeps run.  d_W (may be NULL) + w_off >= 0: warm-start state of the matrix,
 * T10_SVD_WARM_ELEMS(n) elements, zero-initialised by the caller and then left to the kernel -- a ring of two slots
 * used in turn, each keeping the right singular vectors of the call before last on this state, from which the
 * rotations start (any orthogonal start converges; a slowly changing sequence of matrices -- or two interleaved
 * ones, as in iTEBD's alternating bonds -- then needs 2-3 sweeps instead of 10-20); every slot restarts from the
 * identity every 64 uses.  Warm start needs m * n more elements of shared memory.
 * No allocation, no copy, no synchronisation: capturable into a CUDA graph.                                  */
#define T10_SVD_WARM_ELEMS(n) (2 * ((n) * (n) + 1) + 1)
int t10_svd_small_limits(int* max_n, int* max_elems);
int t10_svd_small(int dtype, int64_t nprob, const int64_t* h_prob, const void* d_A, void* d_U, void* d_S,
                  void* d_Vh, void* d_W, int32_t* d_info, t10_stream_t stream);

/* Profiling aid: while d_trace != NULL the fp64 tensor-core kernels record, per tile t of the tile table,
 * d_trace[4t..4t+3] = {SM id, CTA, globaltimer at start, globaltimer at end} (uint64).  NULL turns it off. */
int t10_ggemm_set_trace(void* d_trace);

/* 1 through *h_flag if a tcgen05 launch ever gave up waiting on an MMA barrier (synchronises). */
int t10_ggemm_tf32_status(int* h_flag);

/* Peak-rate probes used by bench.py for the roofline denominators (register-resident loops;
 * mode 1: DMMA, 0: DFMA, 2: DMMA and DFMA warps side by side -- measured 37.0 / 36.5 / 36.8 TF on
 * B200: the two share one FP64 datapath).  Returns achieved TFLOP/s through *tflops.     */
/* DMMA-only probe with a given number of resident warps per SM sub-partition (1 CTA/SM). */
int t10_probe_f64_occupancy(int warps_per_scheduler, int iters, double* tflops, t10_stream_t stream);
int t10_probe_f64_peak(int mode, int iters, double* tflops, t10_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* TOR10_B200_H */
