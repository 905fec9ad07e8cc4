/* hbmpc_b200 -- C-ABI of the B200-native batched share-reconstruction library.
 *
 * This is the drop-in boundary for HoneyBadgerMPC's native math layer.  Each
 * entry point replaces one function the reference exports from its Cython/NTL
 * extension `honeybadgermpc.ntl` (honeybadgermpc/ntl/hbmpc_ntl_helpers.pyx,
 * algorithms in honeybadgermpc/ntl/rsdecode_impl.h); the reference-side binding
 * a maintainer would add is the ctypes shim shown in INTEGRATION.md (ours is
 * honeybadgermpc_b200/ntl/__init__.py).
 *
 * Conventions
 *  - A field element is 4 x uint64_t little-endian limbs (32 bytes), the
 *    canonical residue in [0, p).  Inputs MUST be canonical (the Python shim
 *    reduces like the reference's to_ZZ_p does).  Outputs are canonical.
 *  - Batch arrays are dense row-major: `rows[batch][width]` elements.
 *  - `mem` says where the BATCH buffers (inputs and outputs) live:
 *    HBG_MEM_HOST   -> the library stages them through device memory
 *                      (cudaMemcpyAsync in, kernel, cudaMemcpyAsync out) and
 *                      returns after the result is in the caller's buffer;
 *    HBG_MEM_DEVICE -> device pointers on the context's device; the call only
 *                      enqueues work on the context's stream (see
 *                      hbg_ctx_set_stream / hbg_ctx_synchronize).
 *    Point lists (xs, zs, omega) are always small HOST arrays.
 *  - Every function returns an HBG_* code, never throws, never keeps a caller
 *    pointer after returning.  The caller owns all buffers.
 *  - All batch arithmetic runs in CUDA kernels (sm_100a).  There is no CPU
 *    fallback: without a usable device hbg_ctx_create fails.
 *    Host code only derives the O(n^2) per-point-set constants (Vandermonde
 *    matrices and inverses, twiddles), which are cached in the context.
 *  - A context is bound to one modulus and one device; it may be used from one
 *    thread at a time.
 */
#ifndef HBMPC_B200_H
#define HBMPC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HBG_OK 0
#define HBG_ERR_INVALID 1     /* bad argument (null pointer, size, non-canonical point, even modulus) */
#define HBG_ERR_SINGULAR 2    /* repeated evaluation points: the reference raises InterpolationError
                                 (hbmpc_ntl_helpers.pyx:168-169) */
#define HBG_ERR_CUDA 3        /* a CUDA runtime call failed; see hbg_ctx_last_error */
#define HBG_ERR_UNSUPPORTED 4 /* size outside what the kernels implement */
#define HBG_ERR_NOMEM 5

#define HBG_MEM_HOST 0
#define HBG_MEM_DEVICE 1

typedef struct hbg_ctx hbg_ctx;

/* Library / build identification (also proves the CUDA library is the one loaded). */
const char* hbg_version(void);

/* Replaces NTL's per-call `ZZ_p::init(modulus)` (pyx:107,220,250,...): binds a
 * modulus (odd, 3 <= p < 2^255) to a CUDA device and allocates the stream,
 * scratch space and constant cache. */
int hbg_ctx_create(hbg_ctx** out, const uint64_t modulus[4], int device);
void hbg_ctx_destroy(hbg_ctx* ctx);
const char* hbg_ctx_last_error(const hbg_ctx* ctx);
/* Use the caller's CUDA stream (a cudaStream_t) for all work of this context;
 * NULL restores the context's own stream. */
int hbg_ctx_set_stream(hbg_ctx* ctx, void* cuda_stream);
int hbg_ctx_synchronize(hbg_ctx* ctx);
/* HBG_MEM_HOST calls normally return when the result is in the caller's buffer.
 * With host_async on, the four row-wise batch calls (vandermonde / fft evaluate
 * and interpolate) only ENQUEUE their copies and kernels (a ring of 4 staging
 * slots) and return; the caller keeps its host buffers alive and unmodified and
 * reads results after hbg_ctx_synchronize.  Consecutive calls then overlap on
 * the PCIe link (H2D of one under D2H of the previous).  Host buffers should be
 * pinned for the copies to be asynchronous. */
int hbg_ctx_set_host_async(hbg_ctx* ctx, int on);
/* With host_async on: block until at most `keep` (0..3) of the most recent host-buffer calls of
 * this context are still in flight; the outputs of all earlier calls are then complete in host
 * memory.  keep = 0 is hbg_ctx_synchronize for the host pipeline.  Lets a caller read the
 * results of open i while the transfers of open i+1 run (PCIe is full duplex and the D2H side
 * is the longer one: 46 of the 71 MB of a cfg2 step). */
int hbg_ctx_wait_pending(hbg_ctx* ctx, int keep);
/* Number of kernels this context has launched so far. */
uint64_t hbg_ctx_launch_count(const hbg_ctx* ctx);
/* Name of the dominant kernel of the last batch call (for bench.py / profiles). */
const char* hbg_ctx_last_kernel(const hbg_ctx* ctx);
/* The DFT of hbg_fft_batch_evaluate can run as radix-2 NTT butterflies or as
 * Vandermonde row dot-products on the omega powers (the reference mixes both:
 * a 16-point Vandermonde base case inside the recursion, rsdecode_impl.h:16,
 * :133-136).  Results are bit-identical.  0 = pick the cheaper (default),
 * 1 = dot products, 2 = butterflies, 3 = butterflies through the generic
 * shared-memory kernel even where a register-resident kernel exists (n = 16),
 * 4 = butterflies, n = 16 through the 8-values-in-registers split kernel even
 * where the 4-point-group kernel (d <= 8, the default) applies, 5 = butterflies,
 * n = 16 and d <= 8 through the load-balanced form of the 4-point-group kernel
 * (the two threads of a polynomial hand two values over through shared memory). */
int hbg_ctx_set_fft_path(hbg_ctx* ctx, int path);
/* Which dot-product kernel applies a matrix to the batch (results are
 * bit-identical): 0 = pick by size (default), 1 = matrix read through L1 from
 * global memory, 2 = matrix and TMA-staged input tile in shared memory,
 * 3 = k <= 8 interpolation with the matrix in the constant bank, one row per
 * thread (falls back to 0 where it does not apply), 4 = the same kernel with the
 * carry-free radix-2^29 accumulator instead of the 32-bit-limb carry chains. */
int hbg_ctx_set_matvec_path(hbg_ctx* ctx, int path);

/* vandermonde_batch_evaluate(x, polynomials, modulus), pyx:199-244 +
 * set_vm_matrix rsdecode_impl.h:23-36:
 *   out[b][i] = sum_{l<d} polys[b][l] * xs[i]^l,  b < batch, i < n. */
int hbg_vandermonde_batch_evaluate(hbg_ctx* ctx, const uint64_t* xs, int n,
                                   const uint64_t* polys, size_t batch, int d,
                                   uint64_t* out, int mem);

/* vandermonde_batch_interpolate(x, data_list, modulus), pyx:139-197 +
 * vandermonde_inverse rsdecode_impl.h:97-122:
 *   out[b] = V(xs)^-1 * ys[b]   (k coefficients, NOT stripped).
 * HBG_ERR_SINGULAR if two xs coincide. */
int hbg_vandermonde_batch_interpolate(hbg_ctx* ctx, const uint64_t* xs, int k,
                                      const uint64_t* ys, size_t batch,
                                      uint64_t* out, int mem);

/* fft / partial_fft / fft_batch_evaluate(coeffs, omega, modulus, n, k),
 * pyx:246-316 + fft/_fft rsdecode_impl.h:125-192:
 *   out[b][i] = sum_{j<min(d,n)} polys[b][j] * omega^(i*j),  i < k_out <= n.
 * n must be a power of two and omega a primitive n-th root of unity (the
 * reference's Python callers guarantee this, polynomial.py:117-120; the
 * library checks omega^n == 1 and omega^(n/2) == -1 and returns
 * HBG_ERR_INVALID otherwise). */
int hbg_fft_batch_evaluate(hbg_ctx* ctx, const uint64_t omega[4], int n,
                           const uint64_t* polys, size_t batch, int d, int k_out,
                           uint64_t* out, int mem);

/* fft_interpolate / fft_batch_interpolate(zs, ys_list, omega, modulus, n),
 * pyx:318-381 + fnt_decode_step1/2 rsdecode_impl.h:194-265:
 *   out[b] = the k coefficients of the unique P, deg P < k, with
 *   P(omega^zs[i]) = ys[b][i].   zs are k distinct exponents in [0, n).
 * HBG_ERR_SINGULAR if a z repeats (the reference aborts in inv(0)). */
int hbg_fft_batch_interpolate(hbg_ctx* ctx, const uint64_t omega[4], int n,
                              const int32_t* zs, int k,
                              const uint64_t* ys, size_t batch,
                              uint64_t* out, int mem);

/* hbg_fft_batch_interpolate fused with the all-gather of the multi-GPU path
 * (one rank per GPU, the batch axis sharded): the decoded
 * block of this rank -- rows [rank*batch, (rank+1)*batch) of the gathered
 * [world*batch][k] array -- is written from inside the interpolation kernel
 * straight into EVERY rank's gather buffer.  peer_out[r] is rank r's buffer as
 * mapped into this process (peer / symmetric-memory pointer; peer_out[rank] is
 * the local one).  If multicast_out is non-NULL it is an NVSwitch multicast
 * address bound to all of these buffers and each 16-byte store is issued once
 * (multimem.st) and replicated by the switch.  Device pointers only (ys too);
 * k <= 8.  The caller synchronises the ranks before reading the buffers. */
int hbg_fft_batch_interpolate_allgather(hbg_ctx* ctx, const uint64_t omega[4], int n,
                                        const int32_t* zs, int k,
                                        const uint64_t* ys, size_t batch,
                                        void* const* peer_out, void* multicast_out,
                                        int world, int rank);

/* The all-gather alone: copy an already decoded block (device memory) to byte
 * offset offset_bytes of every rank's gather buffer -- multimem.st through the
 * NVSwitch multicast address when multicast_out is non-NULL, peer stores
 * otherwise -- with at most max_ctas CTAs (0 = 16), so that on a side stream it
 * overlaps the next batch's kernels.  Enqueued on the context's stream. */
int hbg_allgather_block(hbg_ctx* ctx, const void* block, size_t bytes,
                        void* const* peer_out, void* multicast_out,
                        size_t offset_bytes, int world, int max_ctas);

/* The two halves of the hand-over around a fill of slot `slot` that the caller does itself (the
 * fused hbg_fft_batch_interpolate_allgather, whose epilogue stores into every rank's buffer):
 * phase 0, before the first part -- the stream waits until every rank has released the slot;
 * phase 1, after each part -- tells every rank that this rank's part has landed.  Same flag
 * protocol as hbg_allgather_block_signal / _ce, stream-ordered and CUDA-graph capturable. */
int hbg_gather_fence(hbg_ctx* ctx, void* const* flags_peers, int world, int rank, int n_slots, int slot,
                     int parts, int phase);

/* hbg_allgather_block with the slot hand-over on the device (no host-issued barrier).
 * flags_peers[r] is rank r's flag array (uint32, symmetric memory, zero-initialised, at
 * least n_slots * (2 * world + 2) entries) as mapped into this process.  Per slot it holds
 * arrived[world] and released[world] (written by the peers) and, after all slots, the
 * owner's own {sent, waited} counters; every value waited for or published is derived
 * from those device-side counters, so a CUDA graph of these calls can be replayed.
 *   hbg_allgather_block_signal: (first_part != 0: wait until every rank has released the
 *     previous fill of the slot;) copy; publish "one more block of mine has landed" on every
 *     rank.  A fill consists of `parts` blocks per rank.
 *   hbg_gather_wait:    one-warp kernel, returns when every rank's `parts` blocks of the
 *     next unconsumed fill have landed here.
 *   hbg_gather_release: marks that fill consumed and tells every rank.
 * Waits time out after ~2 s (a dead peer must not hang the GPU). */
int hbg_allgather_block_signal(hbg_ctx* ctx, const void* block, size_t bytes,
                               void* const* peer_out, void* multicast_out,
                               size_t offset_bytes, int world, int rank, int max_ctas,
                               void* const* flags_peers, int n_slots, int slot,
                               int parts, int first_part);
int hbg_gather_wait(hbg_ctx* ctx, void* const* flags_peers, int world, int rank,
                    int n_slots, int slot, int parts);
/* The same copy by the TMA unit: max_ctas CTAs of two working threads each stream the block
 * through shared memory (cp.async.bulk global -> shared -> every peer's buffer), so a few SMs keep
 * the NVLink ports busy.  Unicast like the copy engines (world-1 copies leave this GPU, world-1
 * arrive: the lower of the two ingress floors from 4 ranks on -- a multicast store also delivers
 * the sender's own block back to it), a kernel like hbg_allgather_block_signal (one launch, the
 * hand-over inside it). */
int hbg_allgather_block_bulk(hbg_ctx* ctx, const void* block, size_t bytes,
                             void* const* peer_out, size_t offset_bytes, int world, int rank,
                             int max_ctas, void* const* flags_peers, int n_slots, int slot,
                             int parts, int first_part);
/* The same copy on the COPY ENGINES (no SM touches the payload): when first_part != 0 a
 * one-warp kernel first waits for the release of the slot's previous fill; then one
 * cudaMemcpyAsync per peer (peer_out[r] + offset_bytes <- block; the local buffer already
 * holds the block) and a one-warp kernel that publishes the arrival, all stream-ordered on
 * the context's stream.  Unicast: world-1 copies of the block leave this GPU. */
int hbg_allgather_block_ce(hbg_ctx* ctx, const void* block, size_t bytes,
                           void* const* peer_out, size_t offset_bytes, int world, int rank,
                           void* const* flags_peers, int n_slots, int slot,
                           int parts, int first_part);
int hbg_gather_release(hbg_ctx* ctx, void* const* flags_peers, int world, int rank,
                       int n_slots, int slot);

/* gao_interpolate(x, y, k, modulus, ...), pyx:389-439 + gao_interpolate /
 * gao_interpolate_fft / partial_gcd, rsdecode_impl.h:281-405 -- batched: every
 * row of ys is one received word on the SAME m points xs (erasures already
 * removed by the caller, as pyx:399-403 does).  Per row:
 *   status 0: coeffs[b] = the k message coefficients (zero padded),
 *             locator[b][0..loc_len[b]) = the UN-normalised error locator v
 *             (the Bezout cofactor the reference returns; [1] when no error);
 *   status 1: decoding failed (the reference returns (None, None)).
 * loc_stride (elements per locator row) must be >= m - (m+k)/2 + 1. */
int hbg_gao_decode_batch(hbg_ctx* ctx, const uint64_t* xs, int m, int k,
                         const uint64_t* ys, size_t batch,
                         uint64_t* coeffs, uint64_t* locator, int loc_stride,
                         int32_t* loc_len, int32_t* status, int mem);

/* Welch-Berlekamp decode (reed_solomon_wb.py:79-151: solve_system / rref /
 * some_solution) -- batched: every row of ys is one received word on the same
 * m points xs (erasures removed), e_max = (m - (k-1)) / 2 >= 1 as computed at
 * reed_solomon_wb.py:134.  Per row:
 *   status 0: coeffs[b][0..out_len[b]) = P = Q/E with trailing zeros stripped
 *             (the rest of the row is zero);
 *   status 1: ValueError("found no divisors!")  -> caller returns (None, None);
 *   status 2: Exception("No solution")          -> propagates in the reference;
 *   status 3: E came out as the zero polynomial (division by zero).
 * When 2 e_max + k <= m the call may synchronise the stream even for device buffers (it reads
 * the per-word status to pick the words that need the exact elimination kernel). */
int hbg_wb_decode_batch(hbg_ctx* ctx, const uint64_t* xs, int m, int k, int e_max,
                        const uint64_t* ys, size_t batch,
                        uint64_t* coeffs, int32_t* out_len, int32_t* status, int mem);

/* ---- device-resident IncrementalDecoder (reed_solomon.py:232-403) -------------
 * The reference keeps every received column in Python lists, decodes the first
 * degree+1 of them (decoder.decode_batch, :311-313), re-encodes the guess
 * (encoder.encode_batch, :314) and compares each later column with it in a
 * Python loop (:316-319).  These three entry points keep that inner loop on the
 * device: columns live in one buffer colbuf[n][batch] (column i = party i's
 * batch elements, contiguous -- the layout a column arrives in), and only
 * per-column mismatch flags and the final rows travel back. */

/* rows[b][j] = colbuf[idx[j]][b]  (j < k): the (batch x k) matrix of the k chosen
 * columns, row-major, as the interpolation kernels read it.  Device pointers. */
int hbg_columns_to_rows(hbg_ctx* ctx, const uint64_t* colbuf, size_t batch,
                        const int32_t* idx, int k, uint64_t* rows);

/* Fused decode -> re-encode (reed_solomon.py:311-314): for every row the k
 * coefficients of the polynomial through (xs_k[i], ys[b][i]) AND its n evaluations
 * at xs_all, from ONE constant matrix [V(xs_k)^-1 ; V(xs_all) V(xs_k)^-1]:
 *   out[b][0..k)   = coefficients,   out[b][k..k+n) = evaluations.
 * HBG_ERR_SINGULAR if the xs_k repeat. */
int hbg_interpolate_reencode(hbg_ctx* ctx, const uint64_t* xs_k, int k,
                             const uint64_t* xs_all, int n,
                             const uint64_t* ys, size_t batch,
                             uint64_t* out, int mem);

/* Column validation (reed_solomon.py:316-319) for m columns at once:
 *   flags[j] = 1 if colbuf[idx[j]][b] != rows[b][col_offset + idx[j]] for some b,
 *   else 0   (rows has row_width elements per row; for the output of
 *   hbg_interpolate_reencode: row_width = k+n, col_offset = k).
 * rows / colbuf are device pointers; flags_dev is a device int32[m] (always
 * written); if flags_host is non-NULL the flags are also copied there and the
 * call returns after they have arrived. */
int hbg_compare_columns(hbg_ctx* ctx, const uint64_t* rows, int row_width, int col_offset,
                        const uint64_t* colbuf, size_t batch,
                        const int32_t* idx, int m,
                        int32_t* flags_dev, int32_t* flags_host);

/* hbg_wb_decode_batch: 0 = unique-decoding shortcut (words within the decoding radius are
 * decoded by the Gao kernel -- the reference's solver returns the same polynomial for them --
 * and only the others run the exact elimination, which reproduces the reference's failure
 * modes), 1 = exact elimination for every word.  Identical outputs; the tests force both. */
int hbg_ctx_set_wb_path(hbg_ctx* ctx, int path);

/* hbg_fft_batch_interpolate: 0 = automatic (the V^-1 matrix for k <= 128, the
 * NTT-structured path of fnt_decode_step2 above), 1 = matrix, 2 = NTT-structured
 * (HBG_ERR_UNSUPPORTED if the field lacks the root of unity it needs).  All paths
 * return identical bits; the tests force each. */
int hbg_ctx_set_interp_path(hbg_ctx* ctx, int path);

/* Test hook: bound (bytes) of the per-context cache of device constants; when it
 * is exceeded the cache is dropped at the entry of the next call (default 256 MB). */
int hbg_ctx_set_cache_limit(hbg_ctx* ctx, size_t bytes);

/* How the tensor-core kernel writes its results: 0 (default) = one 32-byte store per thread,
 * 1 = transposed through shared memory so that every store instruction writes full 128-byte
 * lines (needs 16-byte aligned outputs).  Identical bits; the tests force both.  Measured on the
 * cfg2 shapes the staged form is the slower one (22.8 vs 18.4 us per step): the two extra
 * shared-memory passes and the 128-thread barriers cost more than the uncoalesced stores. */
int hbg_ctx_set_tc_store(hbg_ctx* ctx, int mode);

/* Upper bound on the CTAs (one per SM) a persistent tensor-core launch of this context uses;
 * 0 = all SMs (default).  A tensor-core CTA owns its SM's shared memory, so two launches of
 * 148 CTAs on two streams run one after the other; two contexts whose limits add up to the SM
 * count run side by side (the NTL reference has no counterpart: it is single-threaded per
 * call, and SetNumThreads, pyx:321, is the nearest knob). */
int hbg_ctx_set_sm_limit(hbg_ctx* ctx, int ctas);

#ifdef __cplusplus
}
#endif
#endif /* HBMPC_B200_H */
