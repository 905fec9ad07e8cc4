// C-ABI of libhbmpc_b200.so (see include/hbmpc_b200.h): context, constant
// cache, staging of host buffers, kernel launches.  Host code here derives the
// O(n^2) per-point-set constants only; every batch element is touched by CUDA
// kernels exclusively.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <string>
#include <type_traits>
#include <unordered_map>
#include <vector>

#include "../../include/hbmpc_b200.h"
#include "host_math.hpp"
#include "kernels.cuh"
#include "robust_kernels.cuh"
#include "tc_kernels.cuh"

using namespace hb;

#define HBG_STR2(x) #x
#define HBG_STR(x) HBG_STR2(x)

namespace {

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
};

struct DevConst {
  void* p = nullptr;
  size_t bytes = 0;
};

}  // namespace

struct hbg_ctx {
  int device = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  FieldParams fp;
  HostField* field = nullptr;
  bool is_bls = false;
  int fft_path = 0;     // 0 auto, 1 matrix, 2 ntt, 3 ntt through the generic smem kernel,
                        // 4 ntt with the register-resident split kernel for n = 16,
                        // 5 ntt, n = 16 with the balanced (hand-over) form of the 4-point-group kernel,
                        // 6 matrix on the tensor cores in the radix-2 split form (where it fits)
  int matvec_path = 0;  // 0 auto, 1 global-memory kernel, 2 shared-memory kernel, 3 small-k kernel,
                        // 4 small-k kernel with the carry-free radix-2^29 arithmetic,
                        // 5 tensor-core kernel (tc_kernels.cuh), 6 never the tensor-core kernel
  unsigned tc_mu = 0;   // floor(2^280 / p) when the tensor-core path serves this modulus, else 0
  unsigned* tc_error = nullptr;  // device word set by a barrier watchdog of tc_apply_kernel
  unsigned* gather_counter = nullptr;  // last-CTA detection of gather_copy_signal_kernel
  cudaStream_t copy_streams[7] = {};   // hbg_allgather_block_ce: one stream per peer, so the copies run
  cudaEvent_t copy_fork = nullptr, copy_join[7] = {};  // on different copy engines concurrently
  int ce_pieces = 1;                   // pieces per peer block in hbg_allgather_block_ce (HBMPC_CE_PIECES)
  int interp_arith = 0; // arithmetic of the small-k kernel when the path is auto / 3
  int interp_path = 0;  // fft_batch_interpolate: 0 auto, 1 V^-1 matrix, 2 NTT-structured (fnt_decode_step2)
  std::string err;
  uint64_t launches = 0;
  const char* last_kernel = "";
  DevBuf in, out, work, work2, fnt_a, fnt_b, wbtmp, wbtmp2;
  int wb_path = 0;      // 0: unique-decoding shortcut + exact kernel for the rest, 1: exact kernel only
  // host-buffer pipeline: H2D / D2H streams and a ring of staging slots, so that with
  // hbg_ctx_set_host_async consecutive calls overlap (call i+1's H2D under call i's D2H)
  cudaStream_t s_in = nullptr, s_out = nullptr;
  struct HostSlot {
    DevBuf in, out;
    cudaEvent_t done = nullptr, ev_in[8] = {}, ev_k[8] = {};
    bool used = false;
  } slots[4];
  unsigned next_slot = 0;
  bool host_async = false;
  int sm_count = 148;
  int sm_limit = 0;     // tensor-core launches use at most this many CTAs (0 = sm_count)
  int tc_store = 0;     // tensor-core epilogue: 0 = 32-byte store per thread, 1 = staged full-line stores (opt-in)
  std::unordered_map<std::string, DevConst> cache;
  std::unordered_map<std::string, std::vector<uint32_t>> host_cache;  // small tables passed by value
  size_t cache_bytes = 0;
  size_t cache_limit = (size_t)256 << 20;
  DevBuf flags;  // hbg_compare_columns: staging for flags_host
};

namespace {

std::mutex g_field_mutex;
struct DeviceFieldState {
  bool valid = false;
  FieldParams fp;
};
DeviceFieldState g_dev_field[64];

int fail(hbg_ctx* c, int code, const std::string& msg) {
  if (c) c->err = msg;
  return code;
}

#define CU(call)                                                                       \
  do {                                                                                 \
    cudaError_t e_ = (call);                                                           \
    if (e_ != cudaSuccess)                                                             \
      return fail(ctx, HBG_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)

// FieldAny kernels read the modulus from the __constant__ bank of this module.
int bind_field(hbg_ctx* ctx, bool needs_constants = false) {
  if (ctx->is_bls && !needs_constants) return HBG_OK;  // FieldBLS: immediates
  std::lock_guard<std::mutex> lk(g_field_mutex);
  DeviceFieldState& st = g_dev_field[ctx->device];
  if (st.valid && memcmp(&st.fp, &ctx->fp, sizeof(FieldParams)) == 0) return HBG_OK;
  CU(cudaDeviceSynchronize());
  CU(cudaMemcpyToSymbol(c_field, &ctx->fp, sizeof(FieldParams)));
  CU(cudaDeviceSynchronize());
  st.valid = true;
  st.fp = ctx->fp;
  return HBG_OK;
}

int ensure(hbg_ctx* ctx, DevBuf& b, size_t bytes) {
  if (bytes <= b.cap) return HBG_OK;
  if (b.p) {
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaFree(b.p));
    b.p = nullptr;
    b.cap = 0;
  }
  size_t cap = bytes + bytes / 4 + 4096;
  cudaError_t e = cudaMalloc(&b.p, cap);
  if (e != cudaSuccess) {
    b.p = nullptr;
    return fail(ctx, HBG_ERR_NOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e));
  }
  b.cap = cap;
  return HBG_OK;
}

// The constant cache is bounded by dropping it wholesale -- but only at the ENTRY of a C-ABI
// call (every batch entry point calls this before its first lookup), never from inside
// get_const: a call may hold several cached pointers at once (Gao: the interpolation matrix
// and g0), and kernels of earlier calls on this stream may still be reading theirs.
int cache_trim(hbg_ctx* ctx) {
  if (ctx->cache_bytes <= ctx->cache_limit) return HBG_OK;
  CU(cudaStreamSynchronize(ctx->stream));
  for (auto& kv : ctx->cache) cudaFree(kv.second.p);
  ctx->cache.clear();
  ctx->host_cache.clear();
  ctx->cache_bytes = 0;
  return HBG_OK;
}

// Cached device constant; `build` fills the host bytes when the key is new (and may add the
// key's host-side twin to host_cache: both maps are only ever cleared together, above).
template <class Build>
int get_const(hbg_ctx* ctx, const std::string& key, const void** out, Build build) {
  auto it = ctx->cache.find(key);
  if (it != ctx->cache.end()) {
    *out = it->second.p;
    return HBG_OK;
  }
  std::vector<uint32_t> host;
  int rc = build(host);
  if (rc != HBG_OK) return rc;
  DevConst dc;
  dc.bytes = host.size() * 4;
  CU(cudaMalloc(&dc.p, dc.bytes ? dc.bytes : 4));
  CU(cudaMemcpyAsync(dc.p, host.data(), dc.bytes, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));  // host vector dies at scope exit
  ctx->cache[key] = dc;
  ctx->cache_bytes += dc.bytes;
  *out = dc.p;
  return HBG_OK;
}

std::string make_key(const char* tag, const void* a, size_t na, const void* b = nullptr,
                     size_t nb = 0, int x = 0, int y = 0) {
  std::string k(tag);
  k.append((const char*)&x, sizeof x);
  k.append((const char*)&y, sizeof y);
  k.append((const char*)a, na);
  if (b) k.append((const char*)b, nb);
  return k;
}

// Points in standard form -> Montgomery form; rejects non-canonical values.
int load_points(hbg_ctx* ctx, const uint64_t* xs, int n, std::vector<Fe>& out) {
  out.resize(n);
  Fe p;
  memcpy(p.w, ctx->fp.p, 32);
  for (int i = 0; i < n; i++) {
    Fe v = fe_from_u64(xs + 4 * i);
    if (fe_cmp(v, p) >= 0) return fail(ctx, HBG_ERR_INVALID, "evaluation point not in [0, p)");
    out[i] = ctx->field->to_mont(v);
  }
  return HBG_OK;
}

// Row-major (rows x cols) matrix -> [col][word][row] interleaved words.
void interleave(const std::vector<Fe>& m, int rows, int cols, std::vector<uint32_t>& out) {
  out.assign((size_t)rows * cols * 8, 0);
  for (int i = 0; i < rows; i++)
    for (int j = 0; j < cols; j++)
      for (int w = 0; w < 8; w++)
        out[((size_t)j * 8 + w) * rows + i] = m[(size_t)i * cols + j].w[w];
}

// Row-wise batch operation with HOST buffers: the batch is cut into up to 8 chunks
// and H2D copy / kernel / D2H copy of consecutive chunks overlap on three streams
// (PCIe is full duplex), so the call costs about max(bytes in, bytes out) / PCIe
// rate instead of their sum.  DEVICE buffers: just the launch.
template <class Launch>
int run_rows(hbg_ctx* ctx, const void* in, size_t in_row, void* out, size_t out_row, size_t batch,
             int mem, Launch launch) {
  if (mem == HBG_MEM_DEVICE) return launch(in, out, batch);
  if (mem != HBG_MEM_HOST) return fail(ctx, HBG_ERR_INVALID, "mem must be HBG_MEM_HOST or HBG_MEM_DEVICE");
  if (!ctx->s_in) {
    CU(cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking));
    for (auto& sl : ctx->slots) {
      CU(cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming));
      for (int i = 0; i < 8; i++) {
        CU(cudaEventCreateWithFlags(&sl.ev_in[i], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&sl.ev_k[i], cudaEventDisableTiming));
      }
    }
  }
  hbg_ctx::HostSlot& sl = ctx->slots[ctx->next_slot++ % 4];
  if (sl.used) CU(cudaEventSynchronize(sl.done));  // the slot's previous call has fully drained
  sl.used = false;
  int rc = ensure(ctx, sl.in, batch * in_row);
  if (rc) return rc;
  rc = ensure(ctx, sl.out, batch * out_row);
  if (rc) return rc;
  size_t chunks = batch * (in_row + out_row) >= ((size_t)4 << 20) ? 8 : 1;
  if (chunks > batch) chunks = 1;
  const size_t per = (batch + chunks - 1) / chunks;
  for (size_t i = 0; i < chunks; i++) {
    const size_t r0 = i * per;
    if (r0 >= batch) break;
    const size_t rows = batch - r0 < per ? batch - r0 : per;
    uint8_t* d_in = (uint8_t*)sl.in.p + r0 * in_row;
    uint8_t* d_out = (uint8_t*)sl.out.p + r0 * out_row;
    if (in_row)
      CU(cudaMemcpyAsync(d_in, (const uint8_t*)in + r0 * in_row, rows * in_row, cudaMemcpyHostToDevice,
                         ctx->s_in));
    CU(cudaEventRecord(sl.ev_in[i], ctx->s_in));
    CU(cudaStreamWaitEvent(ctx->stream, sl.ev_in[i], 0));
    rc = launch(d_in, d_out, rows);
    if (rc) {
      cudaStreamSynchronize(ctx->s_in);
      cudaStreamSynchronize(ctx->stream);
      cudaStreamSynchronize(ctx->s_out);
      return rc;
    }
    CU(cudaEventRecord(sl.ev_k[i], ctx->stream));
    CU(cudaStreamWaitEvent(ctx->s_out, sl.ev_k[i], 0));
    CU(cudaMemcpyAsync((uint8_t*)out + r0 * out_row, d_out, rows * out_row, cudaMemcpyDeviceToHost,
                       ctx->s_out));
  }
  CU(cudaEventRecord(sl.done, ctx->s_out));
  sl.used = true;
  if (!ctx->host_async) {
    CU(cudaEventSynchronize(sl.done));
    sl.used = false;
  }
  return HBG_OK;
}

const size_t kMaxSmem = 226 * 1024;  // 227 KB opt-in limit minus room for static shared memory

// opt in to > 48 KB of dynamic shared memory, once per kernel and device
std::mutex g_smem_mutex;
std::unordered_map<const void*, uint64_t> g_smem_done;  // kernel -> bit mask of devices

template <class K>
int allow_big_smem(hbg_ctx* ctx, K kernel) {
  std::lock_guard<std::mutex> lk(g_smem_mutex);
  uint64_t& mask = g_smem_done[(const void*)kernel];
  if (mask >> ctx->device & 1) return HBG_OK;
  CU(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem));
  mask |= 1ull << ctx->device;
  return HBG_OK;
}

int launch_matvec(hbg_ctx* ctx, const void* mt, int n_out, int d, const void* d_in, int in_stride,
                  void* d_out, int out_stride, size_t batch) {
  MatvecArgs a;
  a.mt = (const uint32_t*)mt;
  a.in = (const uint4*)d_in;
  a.out = (uint4*)d_out;
  a.in_cols = nullptr;
  a.batch = batch;
  a.n_out = n_out;
  a.d = d;
  a.in_stride = in_stride;
  a.out_stride = out_stride;
  a.rows_per_cta = n_out <= 256 ? 256 / n_out : 1;
  a.magic = (unsigned)(((1u << 20) + n_out - 1) / n_out);
  unsigned long long blocks = (batch + a.rows_per_cta - 1) / a.rows_per_cta;
  if (blocks == 0) return HBG_OK;
  if (blocks > 0x7fffffffull) return fail(ctx, HBG_ERR_UNSUPPORTED, "batch too large for one launch");
  int rc = bind_field(ctx);
  if (rc) return rc;
  const size_t m_bytes = (size_t)n_out * d * 32;
  const size_t tile_bytes = (size_t)a.rows_per_cta * d * 32;
  if (ctx->matvec_path != 1 && n_out <= 256 && d >= 1 && in_stride == d && m_bytes <= 100 * 1024 &&
      tile_bytes <= 64 * 1024) {
    size_t smem = m_bytes + tile_bytes;
    if (ctx->is_bls) {
      rc = allow_big_smem(ctx, apply_matrix_smem_kernel<FieldBLS>);
      if (rc) return rc;
      apply_matrix_smem_kernel<FieldBLS><<<(unsigned)blocks, 256, smem, ctx->stream>>>(a);
    } else {
      rc = allow_big_smem(ctx, apply_matrix_smem_kernel<FieldAny>);
      if (rc) return rc;
      apply_matrix_smem_kernel<FieldAny><<<(unsigned)blocks, 256, smem, ctx->stream>>>(a);
    }
    CU(cudaGetLastError());
    ctx->launches++;
    ctx->last_kernel = "apply_matrix_smem_kernel";
    return HBG_OK;
  }
  if (ctx->is_bls)
    apply_matrix_kernel<FieldBLS><<<(unsigned)blocks, 256, 0, ctx->stream>>>(a);
  else
    apply_matrix_kernel<FieldAny><<<(unsigned)blocks, 256, 0, ctx->stream>>>(a);
  CU(cudaGetLastError());
  ctx->launches++;
  ctx->last_kernel = "apply_matrix_kernel";
  return HBG_OK;
}


// ---------------------------------------------------------------------------
// tensor-core path (tc_kernels.cuh): out[b] = M in[b] as an exact u8 GEMM
// ---------------------------------------------------------------------------
struct TcPlan {
  unsigned ob, n_blocks, stages, ew;
  unsigned split = 0, half = 0;  // radix-2 DFT form (tc_kernels.cuh: TcArgs::split)
  unsigned stream = 0, kpad = 0; // constant operand streamed through the stage ring (too large for smem)
  unsigned staged = 0;           // results leave through shared-memory staging, full-line stores (EW = 16)
  size_t b_bytes, smem;
};

const size_t kTcMinBatch = 512;  // below this the IMAD kernels' latency wins (auto mode)

// Shape check: the constant operand (32 n_out x 32 d bytes) stays resident in shared
// memory next to >= 3 stages of 128-row input tiles.
bool tc_plan(int n_out, int d, TcPlan* pl, bool want_staged = false) {
  if (n_out < 1 || d < 1 || d > 1024) return false;
  pl->ob = n_out <= 8 ? (unsigned)n_out : 8u;
  pl->n_blocks = ((unsigned)n_out + pl->ob - 1) / pl->ob;
  pl->b_bytes = (size_t)32 * pl->ob * 32 * d * pl->n_blocks;
  const size_t stage = tc_stage_bytes(32u * d);
  const size_t b_al = (pl->b_bytes + 1023) & ~(size_t)1023;
  size_t room = kMaxSmem - 1024;  // the kernel may skip up to 1008 bytes to align its window
  if (b_al + 2 * stage > room) return false;  // >= 2 stages: one tile in flight behind the MMA
  // staged, full-line stores (opt-in: measured slower than the per-thread stores)
  // if their 32 KB leave room for the stages
  pl->staged = want_staged && b_al + 2 * stage + kTcStoreStaging <= room ? 1u : 0u;
  if (pl->staged) room -= kTcStoreStaging;
  size_t st = (room - b_al) / stage;
  pl->stages = (unsigned)(st > (size_t)kTcMaxStages ? (size_t)kTcMaxStages : st);
  pl->ew = pl->staged || pl->ob % 4 == 0 ? 16 : pl->ob % 3 == 0 ? 12 : 8;
  pl->smem = tc_smem_bytes(32u * d, pl->n_blocks, pl->ob, pl->stages) + 1024 + (pl->staged ? kTcStoreStaging : 0);
  return true;
}

// The DFT matrix (k_out x d, out[i] = sum_j in[j] w^(ij), w of order n) in the radix-2 split form:
// `ob` pairs (i, i + n/2) per block, two half-width accumulator parts per pair.
bool tc_plan_dft(int n, int d, int k_out, TcPlan* pl) {
  if (n < 4 || d < 2 || d > 1024 || k_out < 1) return false;
  const unsigned half = (unsigned)n / 2;
  const unsigned pairs = (unsigned)k_out < half ? (unsigned)k_out : half;
  pl->split = 1;
  pl->half = half;
  pl->ob = pairs < 4 ? pairs : 4;
  pl->n_blocks = (pairs + pl->ob - 1) / pl->ob;
  pl->b_bytes = (size_t)32 * pl->ob * 32 * d * pl->n_blocks;
  const size_t stage = tc_stage_bytes(32u * d);
  const size_t b_al = (pl->b_bytes + 1023) & ~(size_t)1023;
  const size_t room = kMaxSmem - 1024;
  if (b_al + 2 * stage > room) return false;
  size_t st = (room - b_al) / stage;
  pl->stages = (unsigned)(st > (size_t)kTcMaxStages ? (size_t)kTcMaxStages : st);
  pl->ew = pl->ob == 4 ? 16 : pl->ob == 3 ? 12 : 8;
  pl->smem = tc_smem_bytes(32u * d, pl->n_blocks, pl->ob, pl->stages) + 1024;
  return true;
}

// Constant operand of the split form: per block, `ob` pairs; for pair i the E rows hold the bytes
// of w^(2m i) 2^(8a) (K steps = even elements, stored first), the O rows those of
// w^((2m+1) i) 2^(8a) (odd elements, stored after them).
void tc_build_bmat_dft(const HostField& f, const Fe& w_mont, int d, const TcPlan& pl, std::vector<uint32_t>& host) {
  const unsigned NB = 32 * pl.ob, K = 32u * d;
  const unsigned n_even = ((unsigned)d + 1) / 2;
  host.assign(pl.b_bytes / 4, 0);
  uint8_t* b = (uint8_t*)host.data();
  const Fe c256 = f.from_small(256);
  const unsigned pairs = pl.ob * pl.n_blocks;
  for (unsigned i = 0; i < pairs && i < pl.half; i++) {
    const Fe wi = f.pow_u64(w_mont, i);
    const unsigned nb = i / pl.ob, o = i % pl.ob;
    Fe wij = f.one();  // w^(i j)
    for (int j = 0; j < d; j++) {
      const unsigned step = (j & 1) ? n_even + (unsigned)j / 2 : (unsigned)j / 2;
      Fe cur = wij;
      for (unsigned a = 0; a < 32; a++) {
        const Fe s = f.from_mont(cur);
        const uint8_t* bytes = (const uint8_t*)s.w;
        const unsigned kb = step * 32 + a;
        uint8_t* dst = b + (size_t)nb * NB * K + (size_t)(kb / 16) * (NB * 16) + kb % 16;
        for (unsigned cc = 0; cc < 32; cc++) dst[(size_t)(o * 32 + cc) * 16] = bytes[cc];
        cur = f.mul(cur, c256);
      }
      wij = f.mul(wij, wi);
    }
  }
}

// The streamed form: any n_out x d with d <= 1024 (column sums stay below 2^31).
bool tc_plan_stream(int n_out, int d, TcPlan* pl, bool want_staged = false) {
  if (n_out < 1 || d < 1 || d > 1024) return false;
  *pl = TcPlan();
  pl->stream = 1;
  pl->ob = n_out <= 8 ? (unsigned)n_out : 8u;
  pl->n_blocks = ((unsigned)n_out + pl->ob - 1) / pl->ob;
  pl->kpad = ((32u * d + 127) / 128) * 128;
  pl->b_bytes = (size_t)pl->n_blocks * 32 * pl->ob * pl->kpad;
  const size_t stage = tc_stream_stage_bytes(pl->ob);
  const size_t staging = want_staged ? kTcStoreStaging : 0;
  size_t st = (kMaxSmem - 1024 - staging) / stage;
  pl->stages = (unsigned)(st > (size_t)kTcMaxStages ? (size_t)kTcMaxStages : st);
  if (pl->stages < 3) return false;
  pl->staged = want_staged ? 1u : 0u;
  pl->ew = pl->staged || pl->ob % 4 == 0 ? 16 : pl->ob % 3 == 0 ? 12 : 8;
  pl->smem = (size_t)pl->stages * stage + 1024 + staging;
  return true;
}

// Estimated cycles per SM for one 128-row tile: the streamed tensor-core form (the slower of its
// MMA issue, ~300 cycles per 128 x N x 32 MMA measured, and its L2 -> shared-memory traffic at
// ~40 bytes per clock per SM: cfg5's 43 x 43 interpolation measured 0.257 ms per 131 072 rows
// = 72k cycles per tile for 3.2 MB) against the IMAD mat-vec (64 d + 48 IMAD.WIDE per output at 31 per
// clock per SM, ~85 % of the pipe).
double tc_stream_tile_cycles(const TcPlan& pl, int d) {
  const double steps = (double)pl.n_blocks * d;  // MMAs per tile (one K step = one element)
  const double chunks = (double)pl.n_blocks * (pl.kpad / 128);
  const double mma = steps * 300.0, l2 = chunks * (16384.0 + 32.0 * pl.ob * 128.0) / 40.0;
  return mma > l2 ? mma : l2;
}
double imad_matvec_tile_cycles(int n_out, int d) { return 128.0 * n_out * (64.0 * d + 48.0) / 31.0 / 0.85; }

bool tc_wanted(const hbg_ctx* ctx, int n_out, int d, size_t batch, TcPlan* pl) {
  if (!ctx->tc_mu || ctx->matvec_path == 6 || (ctx->matvec_path >= 1 && ctx->matvec_path <= 4)) return false;
  if (!tc_plan(n_out, d, pl, ctx->tc_store == 1)) {
    if (!tc_plan_stream(n_out, d, pl, ctx->tc_store == 1)) return false;
    if (ctx->matvec_path != 5 && tc_stream_tile_cycles(*pl, d) >= imad_matvec_tile_cycles(n_out, d)) return false;
  }
  return ctx->matvec_path == 5 || batch >= kTcMinBatch;
}

// Streamed constant operand: plain row-major [n_blocks * 32 ob][kpad] bytes, row i*32 + c =
// byte c of M[i][j] * 2^(8a) at column j*32 + a.
void tc_build_bplain(const HostField& f, const std::vector<Fe>& m_mont, int n_out, int d, const TcPlan& pl,
                     std::vector<uint32_t>& host) {
  host.assign(pl.b_bytes / 4, 0);
  uint8_t* b = (uint8_t*)host.data();
  const Fe c256 = f.from_small(256);
  for (int i = 0; i < n_out; i++)
    for (int j = 0; j < d; j++) {
      Fe cur = m_mont[(size_t)i * d + j];
      for (unsigned a = 0; a < 32; a++) {
        const Fe s = f.from_mont(cur);
        const uint8_t* bytes = (const uint8_t*)s.w;
        for (unsigned cc = 0; cc < 32; cc++) b[((size_t)i * 32 + cc) * pl.kpad + (size_t)j * 32 + a] = bytes[cc];
        cur = f.mul(cur, c256);
      }
    }
}

// The constant operand of the u8 GEMM for a row-major n_out x d matrix in Montgomery form:
// B[(i,c)][(j,a)] = byte c of (M[i][j] * 2^(8a) mod p), laid out block by block in the UMMA
// canonical K-major form (16-byte K chunks; chunk (n, kc) of a block at kc * NB*16 + n*16).
void tc_build_bmat(const HostField& f, const std::vector<Fe>& m_mont, int n_out, int d, const TcPlan& pl,
                   std::vector<uint32_t>& host) {
  const unsigned NB = 32 * pl.ob, K = 32u * d;
  host.assign(pl.b_bytes / 4, 0);
  uint8_t* b = (uint8_t*)host.data();
  const Fe c256 = f.from_small(256);
  for (int i = 0; i < n_out; i++)
    for (int j = 0; j < d; j++) {
      Fe cur = m_mont[(size_t)i * d + j];
      const unsigned nb = i / pl.ob, o = i % pl.ob;
      for (unsigned a = 0; a < 32; a++) {
        const Fe s = f.from_mont(cur);
        const uint8_t* bytes = (const uint8_t*)s.w;
        const unsigned kb = j * 32 + a;
        uint8_t* dst = b + (size_t)nb * NB * K + (size_t)(kb / 16) * (NB * 16) + kb % 16;
        for (unsigned c = 0; c < 32; c++) dst[(size_t)(o * 32 + c) * 16] = bytes[c];
        cur = f.mul(cur, c256);
      }
    }
}

template <int EW>
int launch_tc_ew(hbg_ctx* ctx, const CUtensorMap& tm, const CUtensorMap& tmb, const TcArgs& a, const TcPlan& pl,
                 unsigned grid) {
  if (pl.stream) {
    int rc = allow_big_smem(ctx, tc_apply_kernel<FieldBLS, EW, true>);
    if (rc) return rc;
    tc_apply_kernel<FieldBLS, EW, true><<<grid, (EW + kTcLoadWarps + 1) * 32, pl.smem, ctx->stream>>>(tm, tmb, a);
    return HBG_OK;
  }
  int rc = allow_big_smem(ctx, tc_apply_kernel<FieldBLS, EW, false>);
  if (rc) return rc;
  tc_apply_kernel<FieldBLS, EW, false><<<grid, (EW + kTcLoadWarps + 1) * 32, pl.smem, ctx->stream>>>(tm, tmb, a);
  return HBG_OK;
}

int launch_tc(hbg_ctx* ctx, const void* d_b, const TcPlan& pl, int n_out, int d, const void* d_in,
              size_t in_pitch, void* d_out, size_t out_pitch, size_t batch,
              const GatherDst* gather = nullptr, size_t gather_row0 = 0) {
  if (batch == 0) return HBG_OK;
  TcArgs a;
  memset(&a, 0, sizeof a);
  if (gather) {
    a.gather_world = (unsigned)gather->world;
    a.gather_mc = (uint8_t*)gather->mc;
    for (int r = 0; r < gather->world; r++) a.gather_peers[r] = (uint8_t*)gather->peers[r];
    a.gather_row0 = gather_row0;
  }
  a.in = (const uint8_t*)d_in;
  a.bmat = (const uint8_t*)d_b;
  a.out = (uint8_t*)d_out;
  a.batch = batch;
  a.K = 32u * d;
  a.n_out = n_out;
  a.ob = pl.ob;
  a.n_blocks = pl.n_blocks;
  a.in_pitch = (unsigned)in_pitch;
  a.out_pitch = (unsigned)out_pitch;
  a.stages = pl.stages;
  a.split = pl.split;
  a.half = pl.half;
  a.mu = ctx->tc_mu;
  a.error = ctx->tc_error;
  const size_t tiles = (batch + 127) / 128;
  const size_t ctas = ctx->sm_limit > 0 && ctx->sm_limit < ctx->sm_count ? ctx->sm_limit : ctx->sm_count;
  const unsigned grid = (unsigned)(tiles < ctas ? tiles : ctas);
  CUtensorMap tm, tmb;
  if (!tc_make_tmap(&tm, d_in, batch, a.K, in_pitch))
    return fail(ctx, HBG_ERR_CUDA, "cuTensorMapEncodeTiled failed (input must be 16-byte aligned)");
  tmb = tm;
  if (pl.stream && !tc_make_tmap_b(&tmb, d_b, (unsigned long long)pl.n_blocks * 32 * pl.ob, pl.kpad, 32 * pl.ob))
    return fail(ctx, HBG_ERR_CUDA, "cuTensorMapEncodeTiled failed for the constant operand");
  // the opt-in staged stores write 16 bytes per thread and need 16-byte aligned destinations
  a.staged = pl.staged && ctx->tc_store == 1 && pl.ew == 16 && !(out_pitch & 15);
  if (gather) {
    if ((uintptr_t)gather->mc & 15) a.staged = 0;
    for (int r = 0; r < gather->world; r++)
      if ((uintptr_t)gather->peers[r] & 15) a.staged = 0;
  } else if ((uintptr_t)d_out & 15) {
    a.staged = 0;
  }
  int rc = pl.ew == 16 ? launch_tc_ew<16>(ctx, tm, tmb, a, pl, grid)
         : pl.ew == 12 ? launch_tc_ew<12>(ctx, tm, tmb, a, pl, grid)
                       : launch_tc_ew<8>(ctx, tm, tmb, a, pl, grid);
  if (rc) return rc;
  CU(cudaGetLastError());
  ctx->launches++;
  ctx->last_kernel = "tc_apply_kernel";
  return HBG_OK;
}

// Cached tensor-core operand for the matrix `gen` produces (row-major, Montgomery form).
template <class Gen>
int tc_const(hbg_ctx* ctx, const std::string& key, int n_out, int d, const TcPlan& pl, const void** d_b,
             Gen gen) {
  return get_const(ctx, key + (pl.stream ? "|tcs" : "|tc"), d_b, [&](std::vector<uint32_t>& host) {
    std::vector<Fe> m;
    int rc = gen(m);
    if (rc) return rc;
    if (pl.stream) tc_build_bplain(*ctx->field, m, n_out, d, pl, host);
    else tc_build_bmat(*ctx->field, m, n_out, d, pl, host);
    return HBG_OK;
  });
}

int ilog2(int n) {
  int l = 0;
  while ((1 << l) < n) l++;
  return l;
}

int check_omega(hbg_ctx* ctx, const uint64_t omega[4], int n, Fe& w_mont) {
  if (n < 1 || (n & (n - 1)) != 0) return fail(ctx, HBG_ERR_INVALID, "fft size must be a power of two");
  std::vector<Fe> w;
  int rc = load_points(ctx, omega, 1, w);
  if (rc) return rc;
  w_mont = w[0];
  const HostField& f = *ctx->field;
  if (n == 1) return HBG_OK;
  Fe half = f.pow_u64(w_mont, (uint64_t)n / 2);
  Fe minus_one = f.neg(f.one());
  if (!fe_eq(half, minus_one))
    return fail(ctx, HBG_ERR_INVALID, "omega is not a primitive n-th root of unity");
  return HBG_OK;
}

int twiddles(hbg_ctx* ctx, const uint64_t omega[4], int n, const void** d_tw,
             const std::vector<uint32_t>** h_tw) {
  std::string key = make_key("tw", omega, 32, nullptr, 0, n);
  int rc = get_const(ctx, key, d_tw, [&](std::vector<uint32_t>& host) {
    Fe w_mont;
    int r = check_omega(ctx, omega, n, w_mont);
    if (r) return r;
    int half = n / 2 > 0 ? n / 2 : 1;
    host.resize((size_t)half * 8);
    Fe acc = ctx->field->one();
    for (int i = 0; i < half; i++) {
      memcpy(&host[(size_t)i * 8], acc.w, 32);
      acc = ctx->field->mul(acc, w_mont);
    }
    if (n <= 16) ctx->host_cache[key] = host;
    return HBG_OK;
  });
  if (rc) return rc;
  auto it = ctx->host_cache.find(key);
  *h_tw = it == ctx->host_cache.end() ? nullptr : &it->second;
  return HBG_OK;
}

template <class F, int D>
int launch_ntt16_t(hbg_ctx* ctx, const Ntt16Args& a) {
  constexpr int POLYS = 64;  // per CTA of 128 threads
  const size_t in_tile = (size_t)POLYS * a.d * 32, out_tile = (size_t)POLYS * ((2 * a.k_out) | 1) * 16;
  const size_t smem = in_tile > out_tile ? in_tile : out_tile;
  int rc = allow_big_smem(ctx, ntt16_split_kernel<F, D, POLYS>);
  if (rc) return rc;
  ntt16_split_kernel<F, D, POLYS><<<(unsigned)((a.batch + POLYS - 1) / POLYS), 2 * POLYS, smem, ctx->stream>>>(a);
  return HBG_OK;
}

// d <= 8: two 4-point groups per thread, results stored from registers (no output tile).
// BAL: the even thread hands the odd thread two of its inputs (rowmath.cuh: ntt16_offload);
// measured not to be a gain, so only `hbg_ctx_set_fft_path(ctx, 5)` selects it.
template <class F, int D, bool BAL>
int launch_ntt16_g4_t(hbg_ctx* ctx, const Ntt16Args& a) {
  constexpr int POLYS = 64;  // per CTA of 128 threads
  // input tile (+ hand-over slots, 4 elements per polynomial, when balancing)
  ntt16_g4_kernel<F, D, POLYS, BAL><<<(unsigned)((a.batch + POLYS - 1) / POLYS), 2 * POLYS,
                                      (size_t)POLYS * a.d * 32 + (BAL ? (size_t)POLYS * 4 * 32 : 0),
                                      ctx->stream>>>(a);
  return HBG_OK;
}

template <class F, bool BAL>
int launch_ntt16_g4_d(hbg_ctx* ctx, const Ntt16Args& a) {
  if (a.d <= 4) return launch_ntt16_g4_t<F, 4, BAL>(ctx, a);
  if (a.d <= 6) return launch_ntt16_g4_t<F, 6, BAL>(ctx, a);
  return launch_ntt16_g4_t<F, 8, BAL>(ctx, a);
}

template <class F>
int launch_ntt16_f(hbg_ctx* ctx, const Ntt16Args& a) {
  if (a.d <= 8 && a.stride == a.d && ctx->fft_path != 4) {
    ctx->last_kernel = "ntt16_g4_kernel";
    return ctx->fft_path == 5 ? launch_ntt16_g4_d<F, true>(ctx, a) : launch_ntt16_g4_d<F, false>(ctx, a);
  }
  ctx->last_kernel = "ntt16_split_kernel";
  if (a.d <= 4) return launch_ntt16_t<F, 4>(ctx, a);
  if (a.d <= 6) return launch_ntt16_t<F, 6>(ctx, a);
  if (a.d <= 8) return launch_ntt16_t<F, 8>(ctx, a);
  if (a.d <= 11) return launch_ntt16_t<F, 11>(ctx, a);
  return launch_ntt16_t<F, 16>(ctx, a);
}

// k x k interpolation with the matrix in the kernel-parameter constant bank
template <class F, int K>
int launch_interp_small_t(hbg_ctx* ctx, const std::vector<uint32_t>& m, const void* d_in, void* d_out,
                          size_t batch, const GatherDst* gather, size_t gather_row0) {
  SmallInterpArgs<K> a;
  memset(&a, 0, sizeof(a));
  a.in = (const uint4*)d_in;
  a.out = (uint4*)d_out;
  a.batch = batch;
  a.gather_row0 = gather_row0;
  if (gather) {
    a.gather = *gather;
  } else {
    memset(&a.gather, 0, sizeof(a.gather));
  }
  // 0: 32-bit-limb lazy accumulator, 1: radix 2^29 (path 4)
  const int arith = ctx->matvec_path == 4 ? 1 : ctx->interp_arith;
  for (int e = 0; e < K * K; e++) {
    Fe v;
    memcpy(v.w, &m[(size_t)e * 8], 32);
    uint32_t* dst = a.mc[e / K][e % K];
    if (arith == 1) {  // limbs of M[i][j] * 2^261 mod p
      Fe r261;
      memcpy(r261.w, ctx->fp.r261, 32);
      to_limbs29(ctx->field->mul(v, r261), dst);
    } else {
      memcpy(dst, v.w, 32);
    }
  }
  // two warps share 32 rows for K >= 4 (each thread K/2 outputs), one thread per row below
  constexpr int ROWS = 64, SPLIT = K >= 4 ? 2 : 1;
  const size_t in_tile = (size_t)ROWS * K * 32, out_tile = (size_t)ROWS * ((2 * K) | 1) * 16;
  const unsigned grid = (unsigned)((batch + ROWS - 1) / ROWS);
  auto go = [&](auto arith_c) {
    constexpr int A = decltype(arith_c)::value;
    if (gather)  // fused all-gather: results go through a shared tile to every rank's buffer
      interp_small_kernel<F, K, ROWS, SPLIT, true, A><<<grid, ROWS * SPLIT, in_tile + out_tile, ctx->stream>>>(a);
    else         // local output: one 256-bit store per element straight from registers
      interp_small_kernel<F, K, ROWS, SPLIT, false, A><<<grid, ROWS * SPLIT, in_tile, ctx->stream>>>(a);
  };
  if (arith == 1) go(std::integral_constant<int, 1>{});
  else go(std::integral_constant<int, 0>{});
  return HBG_OK;
}

template <class F>
int launch_interp_small_f(hbg_ctx* ctx, int k, const std::vector<uint32_t>& m, const void* d_in,
                          void* d_out, size_t batch, const GatherDst* gather = nullptr,
                          size_t gather_row0 = 0) {
  switch (k) {
    case 1: return launch_interp_small_t<F, 1>(ctx, m, d_in, d_out, batch, gather, gather_row0);
    case 2: return launch_interp_small_t<F, 2>(ctx, m, d_in, d_out, batch, gather, gather_row0);
    case 3: return launch_interp_small_t<F, 3>(ctx, m, d_in, d_out, batch, gather, gather_row0);
    case 4: return launch_interp_small_t<F, 4>(ctx, m, d_in, d_out, batch, gather, gather_row0);
    case 5: return launch_interp_small_t<F, 5>(ctx, m, d_in, d_out, batch, gather, gather_row0);
    case 6: return launch_interp_small_t<F, 6>(ctx, m, d_in, d_out, batch, gather, gather_row0);
    case 7: return launch_interp_small_t<F, 7>(ctx, m, d_in, d_out, batch, gather, gather_row0);
    case 8: return launch_interp_small_t<F, 8>(ctx, m, d_in, d_out, batch, gather, gather_row0);
  }
  return HBG_ERR_UNSUPPORTED;
}

// out[b] = M * in[b] for a cached k x k interpolation matrix: the register/constant-bank
// kernel for k <= 8, the generic dot-product kernels otherwise.
int launch_interp(hbg_ctx* ctx, const std::string& key, const void* d_m, int k, const void* d_in,
                  void* d_out, size_t batch) {
  auto it = ctx->host_cache.find(key);
  bool small = k <= 8 && it != ctx->host_cache.end() && (ctx->matvec_path == 0 || ctx->matvec_path >= 3);
  if (!small) return launch_matvec(ctx, d_m, k, k, d_in, k, d_out, k, batch);
  if (batch == 0) return HBG_OK;
  if ((batch + 63) / 64 > 0x7fffffffull) return fail(ctx, HBG_ERR_UNSUPPORTED, "batch too large");
  int rc = bind_field(ctx);
  if (rc) return rc;
  rc = ctx->is_bls ? launch_interp_small_f<FieldBLS>(ctx, k, it->second, d_in, d_out, batch)
                   : launch_interp_small_f<FieldAny>(ctx, k, it->second, d_in, d_out, batch);
  if (rc) return rc;
  CU(cudaGetLastError());
  ctx->launches++;
  ctx->last_kernel = "interp_small_kernel";
  return HBG_OK;
}

int launch_ntt(hbg_ctx* ctx, const void* d_tw, const std::vector<uint32_t>* h_tw, int n,
               const void* d_in, int d, void* d_out, int k_out, size_t batch) {
  int rc = bind_field(ctx);
  if (rc) return rc;
  if (n == 16 && h_tw && ctx->fft_path != 3) {
    Ntt16Args a;
    a.in = (const uint4*)d_in;
    a.out = (uint4*)d_out;
    a.batch = batch;
    a.d = d < 16 ? d : 16;
    a.stride = d;
    a.k_out = k_out;
    memcpy(a.tw, h_tw->data(), 8 * 32);  // omega^0 .. omega^7
    for (int i = 0; i < 8; i++) {         // omega^(8+i) = -omega^i
      Fe w;
      memcpy(w.w, a.tw[i], 32);
      w = ctx->field->neg(w);
      memcpy(a.tw[8 + i], w.w, 32);
    }
    size_t blocks = (batch + 63) / 64;
    if (blocks == 0) return HBG_OK;
    if (blocks > 0x7fffffffull) return fail(ctx, HBG_ERR_UNSUPPORTED, "batch too large for one launch");
    rc = ctx->is_bls ? launch_ntt16_f<FieldBLS>(ctx, a) : launch_ntt16_f<FieldAny>(ctx, a);
    if (rc) return rc;
    CU(cudaGetLastError());
    ctx->launches++;
    return HBG_OK;
  }
  int log_n = ilog2(n);
  if (n <= 1024) {
    NttArgs a;
    a.in = (const uint4*)d_in;
    a.out = (uint4*)d_out;
    a.tw = (const uint4*)d_tw;
    a.batch = batch;
    a.n = n;
    a.log_n = log_n;
    a.d = d;
    a.k_out = k_out;
    int per_cta = n >= 512 ? 1 : 512 / n;
    size_t blocks = (batch + per_cta - 1) / per_cta;
    if (blocks == 0) return HBG_OK;
    if (blocks > 0x7fffffffull) return fail(ctx, HBG_ERR_UNSUPPORTED, "batch too large for one launch");
    size_t smem = (size_t)per_cta * n * 32;
    if (ctx->is_bls)
      ntt_smem_kernel<FieldBLS><<<(unsigned)blocks, 256, smem, ctx->stream>>>(a);
    else
      ntt_smem_kernel<FieldAny><<<(unsigned)blocks, 256, smem, ctx->stream>>>(a);
    CU(cudaGetLastError());
    ctx->launches++;
    ctx->last_kernel = "ntt_smem_kernel";
    return HBG_OK;
  }
  // large n: scatter, one pass per stage, gather
  rc = ensure(ctx, ctx->work, (size_t)batch * n * 32);
  if (rc) return rc;
  NttBigArgs a;
  a.work = (uint4*)ctx->work.p;
  a.in = (const uint4*)d_in;
  a.out = (uint4*)d_out;
  a.tw = (const uint4*)d_tw;
  a.batch = batch;
  a.n = n;
  a.log_n = log_n;
  a.d = d;
  a.k_out = k_out;
  a.stage = 0;
  unsigned long long tot = (unsigned long long)batch * n;
  if ((tot + 255) / 256 > 0x7fffffffull) return fail(ctx, HBG_ERR_UNSUPPORTED, "batch*n too large");
  ntt_big_scatter_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, ctx->stream>>>(a);
  CU(cudaGetLastError());
  ctx->launches++;
  for (int s = 1; s <= log_n; s++) {
    a.stage = s;
    unsigned blocks = (unsigned)((tot / 2 + 255) / 256);
    if (ctx->is_bls)
      ntt_big_stage_kernel<FieldBLS><<<blocks, 256, 0, ctx->stream>>>(a);
    else
      ntt_big_stage_kernel<FieldAny><<<blocks, 256, 0, ctx->stream>>>(a);
    CU(cudaGetLastError());
    ctx->launches++;
  }
  unsigned long long otot = (unsigned long long)batch * k_out;
  if (otot) {
    ntt_big_gather_kernel<<<(unsigned)((otot + 255) / 256), 256, 0, ctx->stream>>>(a);
    CU(cudaGetLastError());
    ctx->launches++;
  }
  ctx->last_kernel = "ntt_big_stage_kernel";
  return HBG_OK;
}

// V(x)^-1 for the k points produced by `points` (Montgomery form), cached by key.
template <class Points>
int interp_matrix(hbg_ctx* ctx, const std::string& key, int k, const void** d_m, Points points,
                  bool montgomery_out = false) {
  return get_const(ctx, key, d_m, [&](std::vector<uint32_t>& host) {
    std::vector<Fe> x_mont, inv;
    int rc = points(x_mont);
    if (rc) return rc;
    if (!vandermonde_inverse(*ctx->field, x_mont, inv))
      return fail(ctx, HBG_ERR_SINGULAR, "evaluation points are not pairwise distinct");
    if (montgomery_out) {  // entries * R: the product with standard-form data is in Montgomery form
      Fe r2;
      memcpy(r2.w, ctx->fp.r2, 32);
      for (Fe& v : inv) v = ctx->field->mul(v, r2);
    }
    interleave(inv, k, k, host);
    if (k <= 8 && !montgomery_out) {
      std::vector<uint32_t> rm((size_t)k * k * 8);
      for (int i = 0; i < k * k; i++) memcpy(&rm[(size_t)i * 8], inv[i].w, 32);
      ctx->host_cache[key] = rm;
    }
    return HBG_OK;
  });
}


// ---------------------------------------------------------------------------
// NTT-structured interpolation (fnt_decode_step1/2, rsdecode_impl.h:194-265).
//   step 1 (host, once per (omega, n, zs), O(k^2)): A = prod (X - x_i), d_i = 1 / A'(x_i),
//           and the transform of A for the truncated product of step 2;
//   step 2 (device, per row): N = sum_i (y_i d_i) X^{z_i}; R = first k+1 values of the size-n
//           transform of N with omega^-1; Q_j = -R_{(j+1) mod n}; P = Q * A mod X^k -- the
//           product as a cyclic convolution of size m >= 2k (forward NTT, pointwise
//           multiply with the cached transform of A scaled by 1/m, inverse NTT).
// No V^-1 (O(k^3) on the host) is ever formed on this path.
// ---------------------------------------------------------------------------
struct FntConst {
  const void* d_scale = nullptr;   // [k] 1/A'(x_i), Montgomery
  const void* d_zs = nullptr;      // [k] int
  const void* d_ahat = nullptr;    // [m] NTT_m(A) / m, Montgomery
  uint64_t omega_inv[4], wm[4], wm_inv[4];
  int m = 0;
};

// a primitive m-th root of unity (m a power of two) in Montgomery form, or false
bool find_root(const HostField& f, const FieldParams& fp, int m, Fe* root) {
  // (p - 1) / m must be exact
  uint32_t pm1[8];
  memcpy(pm1, fp.p, 32);
  pm1[0] -= 1;  // p is odd
  if (m > 1) {
    int bits = 0;
    while ((1 << bits) < m) bits++;
    for (int b = 0; b < bits; b++)
      if ((pm1[b / 32] >> (b % 32)) & 1u) return false;
    for (int b = 0; b < bits; b++) {  // pm1 >>= 1
      for (int w = 0; w < 8; w++) pm1[w] = (pm1[w] >> 1) | (w < 7 ? pm1[w + 1] << 31 : 0);
    }
  }
  Fe e;
  memcpy(e.w, pm1, 32);
  const Fe minus_one = f.neg(f.one());
  for (uint32_t x = 2; x < 200; x++) {
    Fe y = f.pow(f.from_small(x), e);
    if (m == 1) {
      *root = f.one();
      return true;
    }
    if (fe_eq(f.pow_u64(y, (uint64_t)m / 2), minus_one)) {
      *root = y;
      return true;
    }
  }
  return false;
}

void fe_to_std_u64(const HostField& f, const Fe& mont, uint64_t out[4]) {
  Fe s = f.from_mont(mont);
  fe_to_u64(s, out);
}

// host step 1; HBG_ERR_UNSUPPORTED when the field has no root of unity of the order the
// convolution needs (the caller then uses the matrix path)
int fnt_constants(hbg_ctx* ctx, const uint64_t omega[4], int n, const int32_t* zs, int k, FntConst* fc) {
  const HostField& f = *ctx->field;
  int m = 1;
  while (m < 2 * k) m <<= 1;
  fc->m = m;
  Fe w;
  int rc = check_omega(ctx, omega, n, w);
  if (rc) return rc;
  for (int i = 0; i < k; i++)
    if (zs[i] < 0 || zs[i] >= n) return fail(ctx, HBG_ERR_INVALID, "z outside [0, n)");
  Fe wm;
  if (!find_root(f, ctx->fp, m, &wm))
    return fail(ctx, HBG_ERR_UNSUPPORTED, "no root of unity of the order the truncated product needs");
  fe_to_std_u64(f, f.inv(w), fc->omega_inv);
  fe_to_std_u64(f, wm, fc->wm);
  fe_to_std_u64(f, f.inv(wm), fc->wm_inv);
  const std::string key = make_key("fnt", omega, 32, zs, (size_t)k * 4, n, k);
  rc = get_const(ctx, key + "|s", &fc->d_scale, [&](std::vector<uint32_t>& host) {
    std::vector<Fe> xs(k);
    for (int i = 0; i < k; i++) xs[i] = f.pow_u64(w, (uint64_t)zs[i]);
    std::vector<Fe> a = build_from_roots(f, xs);  // k + 1 coefficients
    std::vector<Fe> ad(k);                        // derivative
    for (int i = 0; i < k; i++) ad[i] = f.mul(f.from_small((uint32_t)(i + 1)), a[i + 1]);
    std::vector<Fe> ev(k);
    for (int i = 0; i < k; i++) ev[i] = horner(f, ad, xs[i]);
    if (!f.batch_inv(ev)) return fail(ctx, HBG_ERR_SINGULAR, "repeated z");
    host.resize((size_t)k * 8);
    for (int i = 0; i < k; i++) memcpy(&host[(size_t)i * 8], ev[i].w, 32);
    // the transform of A, scaled by 1/m, goes to a second cache entry
    std::vector<Fe> ah((size_t)m, fe_zero());
    for (int i = 0; i <= k; i++) ah[i] = a[i];
    host_ntt(f, ah, wm);
    const Fe minv = f.inv(f.from_small((uint32_t)m));
    std::vector<uint32_t> h2((size_t)m * 8);
    for (int i = 0; i < m; i++) {
      const Fe v = f.mul(ah[i], minv);
      memcpy(&h2[(size_t)i * 8], v.w, 32);
    }
    ctx->host_cache[key + "|a"] = h2;
    return HBG_OK;
  });
  if (rc) return rc;
  rc = get_const(ctx, key + "|a", &fc->d_ahat, [&](std::vector<uint32_t>& host) {
    auto it = ctx->host_cache.find(key + "|a");
    if (it == ctx->host_cache.end()) return fail(ctx, HBG_ERR_CUDA, "fnt constants out of sync");
    host = it->second;
    ctx->host_cache.erase(it);
    return HBG_OK;
  });
  if (rc) return rc;
  return get_const(ctx, key + "|z", &fc->d_zs, [&](std::vector<uint32_t>& host) {
    host.resize((size_t)k);
    for (int i = 0; i < k; i++) host[i] = (uint32_t)zs[i];
    return HBG_OK;
  });
}

template <class K>
int launch_fnt_kernel(hbg_ctx* ctx, K kernel, const FntArgs& a, unsigned long long elems) {
  unsigned long long blocks = (elems + 255) / 256;
  const unsigned long long cap = (unsigned long long)ctx->sm_count * 16;
  if (blocks > cap) blocks = cap;
  if (blocks == 0) return HBG_OK;
  kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(a);
  CU(cudaGetLastError());
  ctx->launches++;
  return HBG_OK;
}

int launch_ntt(hbg_ctx* ctx, const void* d_tw, const std::vector<uint32_t>* h_tw, int n, const void* d_in,
               int d, void* d_out, int k_out, size_t batch);

int fnt_interpolate(hbg_ctx* ctx, const FntConst& fc, int n, int k, const void* d_ys, size_t batch,
                    void* d_out) {
  const int m = fc.m, kr = k + 1 < n ? k + 1 : n;
  const void *tw_n = nullptr, *tw_m = nullptr, *tw_mi = nullptr;
  const std::vector<uint32_t>*h_n = nullptr, *h_m = nullptr, *h_mi = nullptr;
  int rc = twiddles(ctx, fc.omega_inv, n, &tw_n, &h_n);
  if (rc) return rc;
  rc = twiddles(ctx, fc.wm, m, &tw_m, &h_m);
  if (rc) return rc;
  rc = twiddles(ctx, fc.wm_inv, m, &tw_mi, &h_mi);
  if (rc) return rc;
  rc = bind_field(ctx);
  if (rc) return rc;
  const size_t wide = (size_t)(n > m ? n : m);
  size_t chunk = ((size_t)1 << 30) / (wide * 32);
  if (chunk < 1) chunk = 1;
  if (chunk > batch) chunk = batch;
  rc = ensure(ctx, ctx->fnt_a, chunk * wide * 32);
  if (rc) return rc;
  rc = ensure(ctx, ctx->fnt_b, chunk * (size_t)(kr + k) * 32);
  if (rc) return rc;
  uint4* buf_a = (uint4*)ctx->fnt_a.p;
  uint4* buf_r = (uint4*)ctx->fnt_b.p;
  uint4* buf_q = buf_r + 2 * chunk * (size_t)kr;
  for (size_t r0 = 0; r0 < batch; r0 += chunk) {
    const size_t rows = batch - r0 < chunk ? batch - r0 : chunk;
    FntArgs a;
    memset(&a, 0, sizeof a);
    a.batch = rows;
    a.k = k;
    a.n = n;
    a.m = m;
    // N = scatter of y_i / A'(x_i)
    CU(cudaMemsetAsync(buf_a, 0, rows * (size_t)n * 32, ctx->stream));
    a.in = (const uint4*)d_ys + 2 * r0 * (size_t)k;
    a.out = buf_a;
    a.cst = (const uint4*)fc.d_scale;
    a.zs = (const int*)fc.d_zs;
    rc = ctx->is_bls ? launch_fnt_kernel(ctx, fnt_scale_scatter_kernel<FieldBLS>, a, rows * (size_t)k)
                     : launch_fnt_kernel(ctx, fnt_scale_scatter_kernel<FieldAny>, a, rows * (size_t)k);
    if (rc) return rc;
    // R = first kr values of the transform with omega^-1
    rc = launch_ntt(ctx, tw_n, h_n, n, buf_a, n, buf_r, kr, rows);
    if (rc) return rc;
    // Q_j = -R_{(j+1) mod n}
    a.in = buf_r;
    a.out = buf_q;
    rc = ctx->is_bls ? launch_fnt_kernel(ctx, fnt_shift_negate_kernel<FieldBLS>, a, rows * (size_t)k)
                     : launch_fnt_kernel(ctx, fnt_shift_negate_kernel<FieldAny>, a, rows * (size_t)k);
    if (rc) return rc;
    // P = Q * A mod X^k through a size-m cyclic convolution
    rc = launch_ntt(ctx, tw_m, h_m, m, buf_q, k, buf_a, m, rows);
    if (rc) return rc;
    a.in = buf_a;
    a.out = buf_a;
    a.cst = (const uint4*)fc.d_ahat;
    rc = ctx->is_bls ? launch_fnt_kernel(ctx, fnt_pointwise_kernel<FieldBLS>, a, rows * (size_t)m)
                     : launch_fnt_kernel(ctx, fnt_pointwise_kernel<FieldAny>, a, rows * (size_t)m);
    if (rc) return rc;
    rc = launch_ntt(ctx, tw_mi, h_mi, m, buf_a, m, (uint4*)d_out + 2 * r0 * (size_t)k, k, rows);
    if (rc) return rc;
  }
  ctx->last_kernel = "fnt_decode_step2";
  return HBG_OK;
}

// ---------------------------------------------------------------------------
// robust decoders
// ---------------------------------------------------------------------------
size_t align16(size_t v) { return (v + 15) & ~(size_t)15; }

}  // namespace

extern "C" {

}  // extern "C"

namespace {

// Gao decode of `batch` words already in device memory (all pointers device): constants for the
// point set, the interpolants g1 by one matrix product, the EEA kernel.  Nothing is copied or
// synchronised here.
int gao_device(hbg_ctx* ctx, const uint64_t* xs, int m, int k, const uint4* d_ys, size_t batch, uint4* d_co,
               uint4* d_lo, int loc_stride, int* d_len, int* d_st) {
  const int thr = (m + k) / 2;  // rsdecode_impl.h:338
  // constants: V(x)^-1 scaled into "double Montgomery" form (so that the interpolants come
  // out of apply_matrix in Montgomery form) and g0 = prod (X - x_i)
  const void* d_m = nullptr;
  int rc = interp_matrix(ctx, make_key("gaoinv", xs, (size_t)m * 32, nullptr, 0, m), m, &d_m,
                         [&](std::vector<Fe>& x) { return load_points(ctx, xs, m, x); }, true);
  if (rc) return rc;
  const void* d_g0 = nullptr;
  rc = get_const(ctx, make_key("gaog0", xs, (size_t)m * 32, nullptr, 0, m), &d_g0,
                 [&](std::vector<uint32_t>& host) {
                   std::vector<Fe> x;
                   int r = load_points(ctx, xs, m, x);
                   if (r) return r;
                   std::vector<Fe> g0 = build_from_roots(*ctx->field, x);
                   host.resize(g0.size() * 8);
                   for (size_t i = 0; i < g0.size(); i++) memcpy(&host[i * 8], g0[i].w, 32);
                   return HBG_OK;
                 });
  if (rc) return rc;
  if (batch == 0) return HBG_OK;
  const size_t per_warp = (size_t)8 * (m + 1) * 16;
  int warps = (int)(kMaxSmem / per_warp);
  if (warps < 1) return fail(ctx, HBG_ERR_UNSUPPORTED, "received word too long for shared memory");
  if (warps > 8) warps = 8;
  const size_t ys_b = batch * (size_t)m * 32, lo_b = batch * (size_t)loc_stride * 32;
  rc = ensure(ctx, ctx->work, ys_b);
  if (rc) return rc;
  rc = launch_matvec(ctx, d_m, m, m, d_ys, m, ctx->work.p, m, batch);
  if (rc) return rc;
  rc = bind_field(ctx, true);
  if (rc) return rc;
  GaoArgs a;
  a.g0 = (const uint4*)d_g0;
  a.g1 = (const uint4*)ctx->work.p;
  a.coeffs = d_co;
  a.locator = d_lo;
  a.loc_len = d_len;
  a.status = d_st;
  a.batch = batch;
  a.m = m;
  a.k = k;
  a.thr = thr;
  a.loc_stride = loc_stride;
  a.warps_per_cta = warps;
  size_t blocks = (batch + warps - 1) / warps;
  size_t cap = (size_t)ctx->sm_count * 8;
  if (blocks > cap) blocks = cap;
  size_t smem = per_warp * warps;
  CU(cudaMemsetAsync(d_lo, 0, lo_b, ctx->stream));
  if (ctx->is_bls) {
    rc = allow_big_smem(ctx, gao_kernel<FieldBLS>);
    if (rc) return rc;
    gao_kernel<FieldBLS><<<(unsigned)blocks, 32 * warps, smem, ctx->stream>>>(a);
  } else {
    rc = allow_big_smem(ctx, gao_kernel<FieldAny>);
    if (rc) return rc;
    gao_kernel<FieldAny><<<(unsigned)blocks, 32 * warps, smem, ctx->stream>>>(a);
  }
  CU(cudaGetLastError());
  ctx->launches++;
  ctx->last_kernel = "gao_kernel";
  return HBG_OK;
}

}  // namespace

extern "C" {

int hbg_gao_decode_batch(hbg_ctx* ctx, const uint64_t* xs, int m, int k, const uint64_t* ys,
                         size_t batch, uint64_t* coeffs, uint64_t* locator, int loc_stride,
                         int32_t* loc_len, int32_t* status, int mem) {
  if (!ctx) return HBG_ERR_INVALID;
  if (m < 1 || k < 1 || !xs) return fail(ctx, HBG_ERR_INVALID, "bad size or null points");
  const int thr = (m + k) / 2;  // rsdecode_impl.h:338
  if (loc_stride < m - thr + 1 || loc_stride < 1)
    return fail(ctx, HBG_ERR_INVALID, "loc_stride must be at least m - (m+k)/2 + 1");
  if (mem != HBG_MEM_HOST && mem != HBG_MEM_DEVICE) return fail(ctx, HBG_ERR_INVALID, "bad mem flag");
  CU(cudaSetDevice(ctx->device));
  { int trc = cache_trim(ctx); if (trc) return trc; }
  if (batch && (!ys || !coeffs || !locator || !loc_len || !status))
    return fail(ctx, HBG_ERR_INVALID, "null batch buffer");
  const size_t ys_b = batch * (size_t)m * 32, co_b = batch * (size_t)k * 32;
  const size_t lo_b = batch * (size_t)loc_stride * 32, i_b = align16(batch * 4);
  const uint4* d_ys = (const uint4*)ys;
  uint4 *d_co = (uint4*)coeffs, *d_lo = (uint4*)locator;
  int *d_len = loc_len, *d_st = status;
  int rc;
  if (mem == HBG_MEM_HOST && batch) {
    rc = ensure(ctx, ctx->in, ys_b);
    if (rc) return rc;
    rc = ensure(ctx, ctx->out, co_b + lo_b + 2 * i_b);
    if (rc) return rc;
    CU(cudaMemcpyAsync(ctx->in.p, ys, ys_b, cudaMemcpyHostToDevice, ctx->stream));
    d_ys = (const uint4*)ctx->in.p;
    uint8_t* d_out = (uint8_t*)ctx->out.p;
    d_co = (uint4*)d_out;
    d_lo = (uint4*)(d_out + co_b);
    d_len = (int*)(d_out + co_b + lo_b);
    d_st = (int*)(d_out + co_b + lo_b + i_b);
  }
  rc = gao_device(ctx, xs, m, k, d_ys, batch, d_co, d_lo, loc_stride, d_len, d_st);
  if (rc) return rc;
  if (mem == HBG_MEM_HOST && batch) {
    CU(cudaMemcpyAsync(coeffs, d_co, co_b, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(locator, d_lo, lo_b, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(loc_len, d_len, batch * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(status, d_st, batch * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
  }
  return HBG_OK;
}

}  // extern "C"

namespace {

// Exact Welch-Berlekamp elimination of `batch` words in device memory (no copies, no sync).
int wb_device(hbg_ctx* ctx, const void* d_pw, int pw_stride, int m, int k, int e_max, const uint4* d_ys,
              size_t batch, uint4* d_co, int* d_len, int* d_st) {
  if (batch == 0) return HBG_OK;
  const int nrows = m + 1, max_cols = 2 * e_max + k + 2;
  const size_t scratch = ((size_t)2 * nrows + 2 * max_cols + 2 * m) * 16 +
                         (size_t)((3 * max_cols + 8 + nrows + 3) & ~3) * 4;
  const size_t mat = (size_t)2 * nrows * max_cols * 16;
  const bool in_smem = scratch + mat <= kMaxSmem;
  if (scratch > kMaxSmem) return fail(ctx, HBG_ERR_UNSUPPORTED, "system too large");
  size_t blocks = batch;
  size_t cap = (size_t)ctx->sm_count * (in_smem ? (kMaxSmem / (scratch + mat) > 8 ? 8 : kMaxSmem / (scratch + mat)) : 4);
  if (blocks > cap) blocks = cap;
  WbArgs a;
  a.pw = (const uint4*)d_pw;
  a.ys = d_ys;
  a.coeffs = d_co;
  a.out_len = d_len;
  a.status = d_st;
  a.work = nullptr;
  a.work_stride = mat / 16;
  a.batch = batch;
  a.m = m;
  a.k = k;
  a.e_max = e_max;
  a.pw_stride = pw_stride;
  a.in_smem = in_smem ? 1 : 0;
  int rc;
  if (!in_smem) {
    rc = ensure(ctx, ctx->work2, mat * blocks);
    if (rc) return rc;
    a.work = (uint4*)ctx->work2.p;
  }
  rc = bind_field(ctx, true);
  if (rc) return rc;
  // the kernel writes the coefficients of decoded words only: rows of failed words read as zero
  CU(cudaMemsetAsync(d_co, 0, batch * (size_t)k * 32, ctx->stream));
  size_t smem = scratch + (in_smem ? mat : 0);
  if (ctx->is_bls) {
    rc = allow_big_smem(ctx, wb_kernel<FieldBLS>);
    if (rc) return rc;
    wb_kernel<FieldBLS><<<(unsigned)blocks, kWbThreads, smem, ctx->stream>>>(a);
  } else {
    rc = allow_big_smem(ctx, wb_kernel<FieldAny>);
    if (rc) return rc;
    wb_kernel<FieldAny><<<(unsigned)blocks, kWbThreads, smem, ctx->stream>>>(a);
  }
  CU(cudaGetLastError());
  ctx->launches++;
  ctx->last_kernel = "wb_kernel";
  return HBG_OK;
}

}  // namespace

extern "C" {

int hbg_ctx_set_wb_path(hbg_ctx* ctx, int path) {
  if (!ctx || path < 0 || path > 1) return HBG_ERR_INVALID;
  ctx->wb_path = path;
  return HBG_OK;
}

int hbg_wb_decode_batch(hbg_ctx* ctx, const uint64_t* xs, int m, int k, int e_max,
                        const uint64_t* ys, size_t batch, uint64_t* coeffs, int32_t* out_len,
                        int32_t* status, int mem) {
  if (!ctx) return HBG_ERR_INVALID;
  if (m < 1 || k < 1 || e_max < 1 || !xs) return fail(ctx, HBG_ERR_INVALID, "bad size or null points");
  if (mem != HBG_MEM_HOST && mem != HBG_MEM_DEVICE) return fail(ctx, HBG_ERR_INVALID, "bad mem flag");
  CU(cudaSetDevice(ctx->device));
  { int trc = cache_trim(ctx); if (trc) return trc; }
  const int pw_stride = e_max + k;
  const void* d_pw = nullptr;
  int rc = get_const(ctx, make_key("wbpw", xs, (size_t)m * 32, nullptr, 0, m, pw_stride), &d_pw,
                     [&](std::vector<uint32_t>& host) {
                       std::vector<Fe> x;
                       int r = load_points(ctx, xs, m, x);
                       if (r) return r;
                       host.resize((size_t)m * pw_stride * 8);
                       for (int i = 0; i < m; i++) {
                         Fe acc = ctx->field->one();
                         for (int j = 0; j < pw_stride; j++) {
                           memcpy(&host[((size_t)i * pw_stride + j) * 8], acc.w, 32);
                           acc = ctx->field->mul(acc, x[i]);
                         }
                       }
                       return HBG_OK;
                     });
  if (rc) return rc;
  if (batch == 0) return HBG_OK;
  if (!ys || !coeffs || !out_len || !status) return fail(ctx, HBG_ERR_INVALID, "null batch buffer");

  const size_t ys_b = batch * (size_t)m * 32, co_b = batch * (size_t)k * 32, i_b = align16(batch * 4);
  const uint4* d_ys;
  uint4* d_co;
  int *d_len, *d_st;
  if (mem == HBG_MEM_HOST) {
    rc = ensure(ctx, ctx->in, ys_b);
    if (rc) return rc;
    rc = ensure(ctx, ctx->out, co_b + 2 * i_b);
    if (rc) return rc;
    CU(cudaMemcpyAsync(ctx->in.p, ys, ys_b, cudaMemcpyHostToDevice, ctx->stream));
    d_ys = (const uint4*)ctx->in.p;
    uint8_t* o = (uint8_t*)ctx->out.p;
    d_co = (uint4*)o;
    d_len = (int*)(o + co_b);
    d_st = (int*)(o + co_b + i_b);
  } else {
    d_ys = (const uint4*)ys;
    d_co = (uint4*)coeffs;
    d_len = out_len;
    d_st = status;
  }

  // Unique-decoding shortcut.  With 2 e_max + k <= m (i.e. m - (k-1) odd, as for every n = 3t+1)
  // the word has at most e_max errors iff the Gao decoder succeeds (same capacity), and then
  // EVERY solution (Q, E) of the reference's e = e_max system satisfies Q = P E: Q - P E has degree
  // < e_max + k and vanishes on the >= m - e_max >= e_max + k error-free points.  So the reference
  // returns exactly P = the Gao result, at its first iteration (reed_solomon_wb.py:87-126).  Words the
  // Gao kernel rejects are beyond capacity: the reference ends in one of its failure modes, which the
  // exact elimination kernel reproduces -- it runs on those words only.
  const bool shortcut = ctx->wb_path == 0 && 2 * e_max + k <= m && ctx->sm_count > 0;
  if (!shortcut) {
    rc = wb_device(ctx, d_pw, pw_stride, m, k, e_max, d_ys, batch, d_co, d_len, d_st);
    if (rc) return rc;
  } else {
    const int loc_stride = m - (m + k) / 2 + 1;
    const size_t lo_b = batch * (size_t)loc_stride * 32;
    rc = ensure(ctx, ctx->wbtmp, lo_b + i_b);
    if (rc) return rc;
    uint4* d_lo = (uint4*)ctx->wbtmp.p;
    int* d_ll = (int*)((uint8_t*)ctx->wbtmp.p + lo_b);
    rc = gao_device(ctx, xs, m, k, d_ys, batch, d_co, d_lo, loc_stride, d_ll, d_st);
    if (rc) return rc;
    unsigned long long blocks = (batch + 255) / 256;
    wb_strip_kernel<<<(unsigned)(blocks > 65535 ? 65535 : blocks), 256, 0, ctx->stream>>>(d_co, d_st, d_len, batch, k);
    CU(cudaGetLastError());
    ctx->launches++;
    // which words need the exact kernel?
    std::vector<int> st(batch);
    CU(cudaMemcpyAsync(st.data(), d_st, batch * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    std::vector<int> failed;
    for (size_t i = 0; i < batch; i++)
      if (st[i] != 0) failed.push_back((int)i);
    if (!failed.empty()) {
      const size_t nf = failed.size();
      const size_t f_ys = nf * (size_t)m * 32, f_co = nf * (size_t)k * 32, f_i = align16(nf * 4);
      rc = ensure(ctx, ctx->wbtmp2, f_ys + f_co + 3 * f_i);
      if (rc) return rc;
      uint8_t* t = (uint8_t*)ctx->wbtmp2.p;
      uint4* f_dys = (uint4*)t;
      uint4* f_dco = (uint4*)(t + f_ys);
      int* f_len = (int*)(t + f_ys + f_co);
      int* f_st = (int*)(t + f_ys + f_co + f_i);
      int* f_idx = (int*)(t + f_ys + f_co + 2 * f_i);
      CU(cudaMemcpyAsync(f_idx, failed.data(), nf * 4, cudaMemcpyHostToDevice, ctx->stream));
      rows_gather_kernel<<<(unsigned)((nf * m * 2 + 255) / 256), 256, 0, ctx->stream>>>(d_ys, f_dys, f_idx, nf, m * 2);
      CU(cudaGetLastError());
      rc = wb_device(ctx, d_pw, pw_stride, m, k, e_max, f_dys, nf, f_dco, f_len, f_st);
      if (rc) return rc;
      rows_scatter_kernel<<<(unsigned)((nf * k * 2 + 255) / 256), 256, 0, ctx->stream>>>(f_dco, d_co, f_idx, nf, k * 2,
                                                                                   f_len, f_st, d_len, d_st);
      CU(cudaGetLastError());
      CU(cudaStreamSynchronize(ctx->stream));  // `failed` (host) was the source of an async copy
      ctx->launches += 2;
    }
  }
  if (mem == HBG_MEM_HOST) {
    CU(cudaMemcpyAsync(coeffs, d_co, co_b, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(out_len, d_len, batch * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(status, d_st, batch * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
  }
  return HBG_OK;
}

}  // extern "C"

namespace {
}  // namespace

extern "C" {

const char* hbg_version(void) { return "hbmpc_b200 0.1 (sm_100a, 8x32 Montgomery, CUDA " HBG_STR(CUDART_VERSION) ")"; }

int hbg_ctx_create(hbg_ctx** out, const uint64_t modulus[4], int device) {
  if (!out || !modulus) return HBG_ERR_INVALID;
  *out = nullptr;
  FieldParams fp;
  if (!field_params_init(modulus, &fp)) return HBG_ERR_INVALID;
  if (modulus[3] >> 63) return HBG_ERR_UNSUPPORTED;  // lazy-reduction bounds need p < 2^255
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count || device >= 64)
    return HBG_ERR_CUDA;
  if (cudaSetDevice(device) != cudaSuccess) return HBG_ERR_CUDA;
  hbg_ctx* ctx = new hbg_ctx();
  ctx->device = device;
  ctx->fp = fp;
  ctx->field = new HostField(fp);
  ctx->is_bls = true;
  for (int i = 0; i < 8; i++) ctx->is_bls = ctx->is_bls && fp.p[i] == FieldBLS::p(i);
  if (const char* ar = getenv("HBG_INTERP_ARITH")) {  // experiments: default arithmetic of the small-k kernel
    int v = atoi(ar);
    if (v >= 0 && v <= 1) ctx->interp_arith = v;
  }
  if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete ctx->field;
    delete ctx;
    return HBG_ERR_CUDA;
  }
  ctx->stream = ctx->own_stream;
  cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
  // device words of the gather hand-over kernels (allocated here: the calls that use them
  // must be capturable into CUDA graphs, so they may not allocate)
  if (cudaMalloc(&ctx->gather_counter, 2 * sizeof(unsigned)) != cudaSuccess ||
      cudaMemset(ctx->gather_counter, 0, 2 * sizeof(unsigned)) != cudaSuccess)
    ctx->gather_counter = nullptr;
  if (count > 1) {  // multi-GPU box: the per-peer copy streams of hbg_allgather_block_ce
    if (const char* e = getenv("HBMPC_CE_PIECES")) {
      const int v = atoi(e);
      if (v >= 1 && v <= 7) ctx->ce_pieces = v;
    }
    bool ok = cudaEventCreateWithFlags(&ctx->copy_fork, cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < 7 && ok; i++)
      ok = cudaStreamCreateWithFlags(&ctx->copy_streams[i], cudaStreamNonBlocking) == cudaSuccess &&
           cudaEventCreateWithFlags(&ctx->copy_join[i], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) ctx->copy_fork = nullptr;
  }
  if (ctx->is_bls) {  // the tensor-core path is instantiated for the BLS12-381 scalar field
    ctx->tc_mu = barrett_mu280(fp);
    if (cudaMalloc(&ctx->tc_error, sizeof(unsigned)) != cudaSuccess ||
        cudaMemset(ctx->tc_error, 0, sizeof(unsigned)) != cudaSuccess) {
      ctx->tc_mu = 0;
      ctx->tc_error = nullptr;
    }
  }
  *out = ctx;
  return HBG_OK;
}

void hbg_ctx_destroy(hbg_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (auto& kv : ctx->cache) cudaFree(kv.second.p);
  if (ctx->in.p) cudaFree(ctx->in.p);
  if (ctx->out.p) cudaFree(ctx->out.p);
  if (ctx->work.p) cudaFree(ctx->work.p);
  if (ctx->work2.p) cudaFree(ctx->work2.p);
  if (ctx->tc_error) cudaFree(ctx->tc_error);
  if (ctx->gather_counter) cudaFree(ctx->gather_counter);
  if (ctx->copy_fork) {
    cudaEventDestroy(ctx->copy_fork);
    for (int i = 0; i < 7; i++) {
      if (ctx->copy_streams[i]) {
        cudaStreamSynchronize(ctx->copy_streams[i]);
        cudaStreamDestroy(ctx->copy_streams[i]);
      }
      if (ctx->copy_join[i]) cudaEventDestroy(ctx->copy_join[i]);
    }
  }
  if (ctx->flags.p) cudaFree(ctx->flags.p);
  if (ctx->fnt_a.p) cudaFree(ctx->fnt_a.p);
  if (ctx->fnt_b.p) cudaFree(ctx->fnt_b.p);
  if (ctx->wbtmp.p) cudaFree(ctx->wbtmp.p);
  if (ctx->wbtmp2.p) cudaFree(ctx->wbtmp2.p);
  if (ctx->s_in) {
    cudaStreamSynchronize(ctx->s_in);
    cudaStreamSynchronize(ctx->s_out);
    cudaStreamDestroy(ctx->s_in);
    cudaStreamDestroy(ctx->s_out);
    for (auto& sl : ctx->slots) {
      if (sl.in.p) cudaFree(sl.in.p);
      if (sl.out.p) cudaFree(sl.out.p);
      if (sl.done) cudaEventDestroy(sl.done);
      for (int i = 0; i < 8; i++) {
        if (sl.ev_in[i]) cudaEventDestroy(sl.ev_in[i]);
        if (sl.ev_k[i]) cudaEventDestroy(sl.ev_k[i]);
      }
    }
  }
  cudaStreamDestroy(ctx->own_stream);
  delete ctx->field;
  delete ctx;
}

const char* hbg_ctx_last_error(const hbg_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int hbg_ctx_set_stream(hbg_ctx* ctx, void* cuda_stream) {
  if (!ctx) return HBG_ERR_INVALID;
  ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
  return HBG_OK;
}

int hbg_ctx_synchronize(hbg_ctx* ctx) {
  if (!ctx) return HBG_ERR_INVALID;
  CU(cudaStreamSynchronize(ctx->stream));
  if (ctx->s_out) {
    CU(cudaStreamSynchronize(ctx->s_in));
    CU(cudaStreamSynchronize(ctx->s_out));
    for (auto& sl : ctx->slots) sl.used = false;
  }
  return HBG_OK;
}

int hbg_ctx_wait_pending(hbg_ctx* ctx, int keep) {
  if (!ctx) return HBG_ERR_INVALID;
  if (keep < 0 || keep > 3) return fail(ctx, HBG_ERR_INVALID, "0 <= keep <= 3 (the pipeline has four slots)");
  if (!ctx->s_out) return HBG_OK;  // no host-buffer call yet
  // calls are numbered by next_slot; the newest is next_slot - 1: everything older than the
  // newest `keep` must have drained (D2H complete: its outputs are readable in host memory)
  for (unsigned j = (unsigned)keep; j < 4 && j < ctx->next_slot; j++) {
    hbg_ctx::HostSlot& sl = ctx->slots[(ctx->next_slot - 1 - j) % 4];
    if (sl.used) {
      CU(cudaEventSynchronize(sl.done));
      sl.used = false;
    }
  }
  return HBG_OK;
}

int hbg_ctx_set_host_async(hbg_ctx* ctx, int on) {
  if (!ctx) return HBG_ERR_INVALID;
  if (!on && ctx->host_async) {
    int rc = hbg_ctx_synchronize(ctx);
    if (rc) return rc;
  }
  ctx->host_async = on != 0;
  return HBG_OK;
}

uint64_t hbg_ctx_launch_count(const hbg_ctx* ctx) { return ctx ? ctx->launches : 0; }
const char* hbg_ctx_last_kernel(const hbg_ctx* ctx) { return ctx ? ctx->last_kernel : ""; }

int hbg_ctx_set_matvec_path(hbg_ctx* ctx, int path) {
  if (!ctx || path < 0 || path > 6) return HBG_ERR_INVALID;
  if (path == 5 && !ctx->tc_mu) return fail(ctx, HBG_ERR_UNSUPPORTED, "the tensor-core path serves the BLS12-381 scalar field only");
  ctx->matvec_path = path;
  return HBG_OK;
}

int hbg_ctx_set_tc_store(hbg_ctx* ctx, int mode) {
  if (!ctx || mode < 0 || mode > 1) return HBG_ERR_INVALID;
  ctx->tc_store = mode;
  return HBG_OK;
}

int hbg_ctx_set_interp_path(hbg_ctx* ctx, int path) {
  if (!ctx || path < 0 || path > 2) return HBG_ERR_INVALID;
  ctx->interp_path = path;
  return HBG_OK;
}

int hbg_ctx_set_fft_path(hbg_ctx* ctx, int path) {
  if (!ctx || path < 0 || path > 6) return HBG_ERR_INVALID;
  ctx->fft_path = path;
  return HBG_OK;
}

int hbg_vandermonde_batch_evaluate(hbg_ctx* ctx, const uint64_t* xs, int n, const uint64_t* polys,
                                   size_t batch, int d, uint64_t* out, int mem) {
  if (!ctx) return HBG_ERR_INVALID;
  if (n < 0 || d < 0 || (n && !xs)) return fail(ctx, HBG_ERR_INVALID, "bad size or null points");
  if (batch == 0 || n == 0) return HBG_OK;
  if (!out || (d && !polys)) return fail(ctx, HBG_ERR_INVALID, "null batch buffer");
  CU(cudaSetDevice(ctx->device));
  { int trc = cache_trim(ctx); if (trc) return trc; }
  // V[i][j] = x_i^j, row-major, Montgomery form (set_vm_matrix, rsdecode_impl.h:23-36)
  auto gen = [&](std::vector<Fe>& m) {
    std::vector<Fe> x;
    int r = load_points(ctx, xs, n, x);
    if (r) return r;
    m.assign((size_t)n * (d ? d : 1), fe_zero());
    for (int i = 0; i < n; i++) {
      Fe acc = ctx->field->one();
      for (int j = 0; j < d; j++) {
        m[(size_t)i * d + j] = acc;
        acc = ctx->field->mul(acc, x[i]);
      }
    }
    return HBG_OK;
  };
  const std::string key = make_key("vdm", xs, (size_t)n * 32, nullptr, 0, n, d);
  TcPlan pl;
  const bool tc = d > 0 && tc_wanted(ctx, n, d, batch, &pl);
  const void* d_m = nullptr;
  int rc = tc ? tc_const(ctx, key, n, d, pl, &d_m, gen)
              : get_const(ctx, key, &d_m, [&](std::vector<uint32_t>& host) {
                  std::vector<Fe> m;
                  int r = gen(m);
                  if (r) return r;
                  interleave(m, n, d, host);
                  return HBG_OK;
                });
  if (rc) return rc;
  return run_rows(ctx, polys, (size_t)d * 32, out, (size_t)n * 32, batch, mem,
                  [&](const void* di, void* dout, size_t rows) {
                    if (tc) return launch_tc(ctx, d_m, pl, n, d, di, (size_t)d * 32, dout, (size_t)n * 32, rows);
                    return launch_matvec(ctx, d_m, n, d, di, d, dout, n, rows);
                  });
}

int hbg_vandermonde_batch_interpolate(hbg_ctx* ctx, const uint64_t* xs, int k, const uint64_t* ys,
                                      size_t batch, uint64_t* out, int mem) {
  if (!ctx) return HBG_ERR_INVALID;
  if (k < 0 || (k && !xs)) return fail(ctx, HBG_ERR_INVALID, "bad size or null points");
  CU(cudaSetDevice(ctx->device));
  { int trc = cache_trim(ctx); if (trc) return trc; }
  const void* d_m = nullptr;
  // the singularity check must run even for an empty batch (pyx:167-169)
  const std::string key = make_key("vinv", xs, (size_t)k * 32, nullptr, 0, k);
  int rc = interp_matrix(ctx, key, k, &d_m,
                         [&](std::vector<Fe>& x) { return load_points(ctx, xs, k, x); });
  if (rc) return rc;
  if (batch == 0 || k == 0) return HBG_OK;
  if (!out || !ys) return fail(ctx, HBG_ERR_INVALID, "null batch buffer");
  TcPlan pl;
  const bool tc = tc_wanted(ctx, k, k, batch, &pl);
  const void* d_b = nullptr;
  if (tc) {
    rc = tc_const(ctx, key, k, k, pl, &d_b, [&](std::vector<Fe>& inv) {
      std::vector<Fe> x;
      int r = load_points(ctx, xs, k, x);
      if (r) return r;
      return vandermonde_inverse(*ctx->field, x, inv) ? HBG_OK : HBG_ERR_SINGULAR;
    });
    if (rc) return rc;
  }
  return run_rows(ctx, ys, (size_t)k * 32, out, (size_t)k * 32, batch, mem,
                  [&](const void* di, void* dout, size_t rows) {
                    if (tc) return launch_tc(ctx, d_b, pl, k, k, di, (size_t)k * 32, dout, (size_t)k * 32, rows);
                    return launch_interp(ctx, key, d_m, k, di, dout, rows);
                  });
}

int hbg_allgather_block(hbg_ctx* ctx, const void* block, size_t bytes, void* const* peer_out,
                        void* multicast_out, size_t offset_bytes, int world, int max_ctas) {
  if (!ctx) return HBG_ERR_INVALID;
  if (!block || !peer_out || world < 1 || world > 8 || (bytes & 15) || (offset_bytes & 15))
    return fail(ctx, HBG_ERR_INVALID, "bad argument (sizes and offsets must be multiples of 16)");
  if (bytes == 0) return HBG_OK;
  CU(cudaSetDevice(ctx->device));
  { int trc = cache_trim(ctx); if (trc) return trc; }
  GatherDst g;
  memset(&g, 0, sizeof(g));
  g.world = world;
  g.mc = (uint4*)multicast_out;
  for (int r = 0; r < world; r++) {
    if (!peer_out[r]) return fail(ctx, HBG_ERR_INVALID, "null peer pointer");
    g.peers[r] = (uint4*)peer_out[r];
  }
  unsigned long long chunks = bytes / 16;
  unsigned long long want = (chunks + 255) / 256;
  unsigned ctas = (unsigned)(max_ctas > 0 ? max_ctas : 16);
  if (ctas > want) ctas = (unsigned)want;
  gather_copy_kernel<<<ctas, 256, 0, ctx->stream>>>((const uint4*)block, g, offset_bytes / 16, chunks);
  CU(cudaGetLastError());
  ctx->launches++;
  ctx->last_kernel = "gather_copy_kernel";
  return HBG_OK;
}

namespace {
int gather_signal(hbg_ctx* ctx, void* const* flags_peers, int world, int rank, int n_slots, int slot,
                  int parts, GatherSignal* s) {
  if (!flags_peers || world < 1 || world > 8 || rank < 0 || rank >= world || slot < 0 || slot >= n_slots ||
      parts < 1)
    return fail(ctx, HBG_ERR_INVALID, "bad gather signal argument");
  memset(s, 0, sizeof *s);
  for (int r = 0; r < world; r++) {
    if (!flags_peers[r]) return fail(ctx, HBG_ERR_INVALID, "null flag pointer");
    s->flags[r] = (unsigned*)flags_peers[r];
  }
  if (!ctx->gather_counter) return fail(ctx, HBG_ERR_NOMEM, "gather counter was not allocated");
  s->counter = ctx->gather_counter;
  s->error = ctx->gather_counter + 1;
  s->slot_base = (unsigned)slot * 2u * (unsigned)world;
  s->local_base = (unsigned)n_slots * 2u * (unsigned)world + 2u * (unsigned)slot;
  s->parts = (unsigned)parts;
  s->world = world;
  s->rank = rank;
  return HBG_OK;
}
}  // namespace

int hbg_allgather_block_signal(hbg_ctx* ctx, const void* block, size_t bytes, void* const* peer_out,
                               void* multicast_out, size_t offset_bytes, int world, int rank, int max_ctas,
                               void* const* flags_peers, int n_slots, int slot, int parts, int first_part) {
  if (!ctx) return HBG_ERR_INVALID;
  if (!block || !peer_out || world < 1 || world > 8 || (bytes & 15) || (offset_bytes & 15) || bytes == 0)
    return fail(ctx, HBG_ERR_INVALID, "bad argument (sizes and offsets must be non-zero multiples of 16)");
  CU(cudaSetDevice(ctx->device));
  GatherSignal s;
  int rc = gather_signal(ctx, flags_peers, world, rank, n_slots, slot, parts, &s);
  if (rc) return rc;
  s.first_part = first_part;
  GatherDst g;
  memset(&g, 0, sizeof(g));
  g.world = world;
  g.mc = (uint4*)multicast_out;
  for (int r = 0; r < world; r++) {
    if (!peer_out[r]) return fail(ctx, HBG_ERR_INVALID, "null peer pointer");
    g.peers[r] = (uint4*)peer_out[r];
  }
  unsigned long long chunks = bytes / 16;
  unsigned long long want = (chunks + 255) / 256;
  unsigned ctas = (unsigned)(max_ctas > 0 ? max_ctas : 16);
  if (ctas > want) ctas = (unsigned)want;
  gather_copy_signal_kernel<<<ctas, 256, 0, ctx->stream>>>((const uint4*)block, g, offset_bytes / 16, chunks, s);
  CU(cudaGetLastError());
  ctx->launches++;
  ctx->last_kernel = "gather_copy_signal_kernel";
  return HBG_OK;
}

int hbg_allgather_block_bulk(hbg_ctx* ctx, const void* block, size_t bytes, void* const* peer_out,
                             size_t offset_bytes, int world, int rank, int max_ctas, void* const* flags_peers,
                             int n_slots, int slot, int parts, int first_part) {
  if (!ctx) return HBG_ERR_INVALID;
  if (!block || !peer_out || world < 1 || world > 8 || (bytes & 15) || (offset_bytes & 15) || bytes == 0 ||
      ((uintptr_t)block & 15))
    return fail(ctx, HBG_ERR_INVALID, "bad argument (pointers, sizes and offsets must be non-zero multiples of 16)");
  CU(cudaSetDevice(ctx->device));
  GatherSignal s;
  int rc = gather_signal(ctx, flags_peers, world, rank, n_slots, slot, parts, &s);
  if (rc) return rc;
  s.first_part = first_part;
  GatherDst g;
  memset(&g, 0, sizeof(g));
  g.world = world;
  for (int r = 0; r < world; r++) {
    if (!peer_out[r] || ((uintptr_t)peer_out[r] & 15)) return fail(ctx, HBG_ERR_INVALID, "null or unaligned peer pointer");
    g.peers[r] = (uint4*)peer_out[r];
  }
  const size_t smem = (size_t)kGbStages * kGbChunk;
  rc = allow_big_smem(ctx, gather_bulk_signal_kernel);
  if (rc) return rc;
  unsigned long long want = (bytes + kGbChunk - 1) / kGbChunk;
  unsigned ctas = (unsigned)(max_ctas > 0 ? max_ctas : 16);
  if (ctas > want) ctas = (unsigned)want;
  gather_bulk_signal_kernel<<<ctas, 64, smem, ctx->stream>>>((const uint8_t*)block, g, offset_bytes, bytes, s);
  CU(cudaGetLastError());
  ctx->launches++;
  ctx->last_kernel = "gather_bulk_signal_kernel";
  return HBG_OK;
}

int hbg_allgather_block_ce(hbg_ctx* ctx, const void* block, size_t bytes, void* const* peer_out,
                           size_t offset_bytes, int world, int rank, void* const* flags_peers, int n_slots,
                           int slot, int parts, int first_part) {
  if (!ctx) return HBG_ERR_INVALID;
  if (!block || !peer_out || world < 1 || world > 8 || bytes == 0)
    return fail(ctx, HBG_ERR_INVALID, "bad argument");
  CU(cudaSetDevice(ctx->device));
  GatherSignal s;
  int rc = gather_signal(ctx, flags_peers, world, rank, n_slots, slot, parts, &s);
  if (rc) return rc;
  if (first_part) {
    gather_wait_released_kernel<<<1, 32, 0, ctx->stream>>>(s);
    CU(cudaGetLastError());
    ctx->launches++;
  }
  // fork: one copy per peer, each on its own stream (= its own copy engine), then join
  if (world > 1 && !ctx->copy_fork) return fail(ctx, HBG_ERR_CUDA, "copy streams were not created");
  if (world > 1) CU(cudaEventRecord(ctx->copy_fork, ctx->stream));
  // One copy engine moves ~570 GB/s of a 770 GB/s link (22 us for 12.6 MB).  Cutting a peer's block
  // into pieces on several streams (= engines; HBMPC_CE_PIECES, up to 7 streams in all, pieces of at
  // least 1 MB) was measured SLOWER at 2 ranks -- 29.2 / 31.6 / 34.7 / 38.8 us per cfg2 step for
  // 1 / 2 / 3 / 4 pieces (profiles/r2g_ce_pieces_n2.jsonl): every extra copy is another engine
  // switch between kernels -- so the default is one copy per peer.
  int pieces = world > 1 ? 7 / (world - 1) : 1;
  if (pieces > ctx->ce_pieces) pieces = ctx->ce_pieces;
  while (pieces > 1 && bytes / (size_t)pieces < ((size_t)1 << 20)) pieces--;
  if (pieces < 1) pieces = 1;
  const size_t per = ((bytes / (size_t)pieces) + 255) & ~(size_t)255;
  int sidx = 0;
  for (int i = 1; i < world; i++) {
    const int r = (rank + i) % world;  // staggered: at any moment the ranks target different peers
    if (!peer_out[r]) return fail(ctx, HBG_ERR_INVALID, "null peer pointer");
    for (int pc = 0; pc < pieces; pc++, sidx++) {
      const size_t o = (size_t)pc * per;
      if (o >= bytes) break;
      const size_t len = bytes - o < per || pc == pieces - 1 ? bytes - o : per;
      cudaStream_t cs = ctx->copy_streams[sidx];
      CU(cudaStreamWaitEvent(cs, ctx->copy_fork, 0));
      CU(cudaMemcpyAsync((uint8_t*)peer_out[r] + offset_bytes + o, (const uint8_t*)block + o, len,
                         cudaMemcpyDeviceToDevice, cs));
      CU(cudaEventRecord(ctx->copy_join[sidx], cs));
      CU(cudaStreamWaitEvent(ctx->stream, ctx->copy_join[sidx], 0));
    }
  }
  gather_signal_arrived_kernel<<<1, 32, 0, ctx->stream>>>(s);
  CU(cudaGetLastError());
  ctx->launches++;
  ctx->last_kernel = "gather_signal_arrived_kernel";
  return HBG_OK;
}

int hbg_gather_fence(hbg_ctx* ctx, void* const* flags_peers, int world, int rank, int n_slots, int slot,
                     int parts, int phase) {
  if (!ctx) return HBG_ERR_INVALID;
  if (phase != 0 && phase != 1) return fail(ctx, HBG_ERR_INVALID, "phase is 0 (before the fill) or 1 (after it)");
  CU(cudaSetDevice(ctx->device));
  GatherSignal s;
  int rc = gather_signal(ctx, flags_peers, world, rank, n_slots, slot, parts, &s);
  if (rc) return rc;
  if (phase == 0)
    gather_wait_released_kernel<<<1, 32, 0, ctx->stream>>>(s);
  else
    gather_signal_arrived_kernel<<<1, 32, 0, ctx->stream>>>(s);
  CU(cudaGetLastError());
  ctx->launches++;
  return HBG_OK;
}

int hbg_gather_wait(hbg_ctx* ctx, void* const* flags_peers, int world, int rank, int n_slots, int slot,
                    int parts) {
  if (!ctx) return HBG_ERR_INVALID;
  CU(cudaSetDevice(ctx->device));
  GatherSignal s;
  int rc = gather_signal(ctx, flags_peers, world, rank, n_slots, slot, parts, &s);
  if (rc) return rc;
  gather_wait_kernel<<<1, 32, 0, ctx->stream>>>(s);
  CU(cudaGetLastError());
  ctx->launches++;
  return HBG_OK;
}

int hbg_gather_release(hbg_ctx* ctx, void* const* flags_peers, int world, int rank, int n_slots, int slot) {
  if (!ctx) return HBG_ERR_INVALID;
  CU(cudaSetDevice(ctx->device));
  GatherSignal s;
  int rc = gather_signal(ctx, flags_peers, world, rank, n_slots, slot, 1, &s);
  if (rc) return rc;
  gather_release_kernel<<<1, 32, 0, ctx->stream>>>(s);
  CU(cudaGetLastError());
  ctx->launches++;
  return HBG_OK;
}

int hbg_fft_batch_interpolate_allgather(hbg_ctx* ctx, const uint64_t omega[4], int n, const int32_t* zs,
                                        int k, const uint64_t* ys, size_t batch, void* const* peer_out,
                                        void* multicast_out, int world, int rank) {
  if (!ctx) return HBG_ERR_INVALID;
  if (!omega || k < 1 || !zs || !ys || !peer_out || world < 1 || world > 8 || rank < 0 || rank >= world)
    return fail(ctx, HBG_ERR_INVALID, "bad argument");
  TcPlan pl;
  const bool tc = tc_wanted(ctx, k, k, batch, &pl);
  if (k > 8 && !tc)
    return fail(ctx, HBG_ERR_UNSUPPORTED,
                "fused all-gather: k <= 8, or a shape the tensor-core kernel serves (BLS12-381 field, "
                "constant operand resident in shared memory)");
  CU(cudaSetDevice(ctx->device));
  { int trc = cache_trim(ctx); if (trc) return trc; }
  const void* d_m = nullptr;
  const std::string key = make_key("finv", omega, 32, zs, (size_t)k * 4, n, k);
  int rc = interp_matrix(ctx, key, k, &d_m,
                         [&](std::vector<Fe>& x) {
                           Fe w;
                           int r = check_omega(ctx, omega, n, w);
                           if (r) return r;
                           x.resize(k);
                           for (int i = 0; i < k; i++) {
                             if (zs[i] < 0 || zs[i] >= n)
                               return fail(ctx, HBG_ERR_INVALID, "z outside [0, n)");
                             x[i] = ctx->field->pow_u64(w, (uint64_t)zs[i]);
                           }
                           return HBG_OK;
                         });
  if (rc) return rc;
  if (batch == 0) return HBG_OK;
  GatherDst g;
  memset(&g, 0, sizeof(g));
  g.world = world;
  g.mc = (uint4*)multicast_out;
  for (int r = 0; r < world; r++) {
    if (!peer_out[r]) return fail(ctx, HBG_ERR_INVALID, "null peer pointer");
    g.peers[r] = (uint4*)peer_out[r];
  }
  if (tc) {  // tensor-core kernel, results stored from its epilogue into every rank's buffer
    const void* d_b = nullptr;
    rc = tc_const(ctx, key, k, k, pl, &d_b, [&](std::vector<Fe>& inv) {
      Fe w;
      int r = check_omega(ctx, omega, n, w);
      if (r) return r;
      std::vector<Fe> x(k);
      for (int i = 0; i < k; i++) x[i] = ctx->field->pow_u64(w, (uint64_t)zs[i]);
      return vandermonde_inverse(*ctx->field, x, inv) ? HBG_OK : HBG_ERR_SINGULAR;
    });
    if (rc) return rc;
    return launch_tc(ctx, d_b, pl, k, k, ys, (size_t)k * 32, nullptr, (size_t)k * 32, batch, &g,
                     (size_t)rank * batch);
  }
  auto it = ctx->host_cache.find(key);
  if (it == ctx->host_cache.end()) return fail(ctx, HBG_ERR_UNSUPPORTED, "matrix not cached on host");
  rc = bind_field(ctx);
  if (rc) return rc;
  rc = ctx->is_bls ? launch_interp_small_f<FieldBLS>(ctx, k, it->second, ys, nullptr, batch, &g,
                                                     (size_t)rank * batch)
                   : launch_interp_small_f<FieldAny>(ctx, k, it->second, ys, nullptr, batch, &g,
                                                     (size_t)rank * batch);
  if (rc) return rc;
  CU(cudaGetLastError());
  ctx->launches++;
  ctx->last_kernel = "interp_small_kernel";
  return HBG_OK;
}

int hbg_fft_batch_evaluate(hbg_ctx* ctx, const uint64_t omega[4], int n, const uint64_t* polys,
                           size_t batch, int d, int k_out, uint64_t* out, int mem) {
  if (!ctx) return HBG_ERR_INVALID;
  if (!omega || d < 0 || k_out < 0 || k_out > n)
    return fail(ctx, HBG_ERR_INVALID, "bad size or null omega");
  CU(cudaSetDevice(ctx->device));
  { int trc = cache_trim(ctx); if (trc) return trc; }
  if (n < 1 || (n & (n - 1)) != 0) return fail(ctx, HBG_ERR_INVALID, "fft size must be a power of two");
  int rc;
  if (batch == 0 || k_out == 0) {
    Fe w;
    return check_omega(ctx, omega, n, w);
  }
  if (!out || (d && !polys)) return fail(ctx, HBG_ERR_INVALID, "null batch buffer");
  const int d_eff = d < n ? d : n;
  // cost in IMAD.WIDE per polynomial: dot products vs butterflies
  double cost_mat = (double)k_out * d_eff * 64 + 64.0 * k_out;
  double cost_ntt = (double)(n / 2) * (ilog2(n) > 1 ? ilog2(n) - 1 : 0) * 120 + 1;
  bool use_matrix = n < 2 || cost_mat <= cost_ntt;
  if (ctx->fft_path == 1 && (size_t)k_out * d_eff <= (1u << 22)) use_matrix = true;
  if (ctx->fft_path >= 2 && ctx->fft_path != 6 && n >= 2) use_matrix = false;
  if ((size_t)k_out * d_eff > (1u << 22)) use_matrix = false;
  // the matrix form on the tensor cores beats both whenever its constant operand fits
  TcPlan pl;
  bool tc = d_eff > 0 && (ctx->fft_path <= 1 || n < 2) && tc_wanted(ctx, k_out, d_eff, batch, &pl);
  // a streamed operand competes with the butterflies (IMAD.WIDE at ~56 % of the pipe for the
  // shared-memory NTT): automatic mode keeps the NTT when it is estimated faster
  if (tc && pl.stream && ctx->fft_path == 0 && ctx->matvec_path != 5 && n >= 2 &&
      tc_stream_tile_cycles(pl, d_eff) >= 128.0 * cost_ntt / 31.0 / 0.56)
    tc = false;
  // The radix-2 split form (half the multiply-accumulates, half the constant operand) is taken
  // when asked for (fft path 6), or when only ITS operand fits shared memory.  It is not the
  // default: measured, an MMA of this kernel costs about the same at N = 128 and N = 256 (operand
  // fetch of the 128 x 32-byte A slice dominates), so halving N buys nothing and the butterfly
  // epilogue costs a little (encode of 65 536 x 6 -> 16: 26 us split, 21 us plain).
  const bool split_ok = d_eff > 0 && ctx->tc_mu && ctx->matvec_path != 6 &&
                        !(ctx->matvec_path >= 1 && ctx->matvec_path <= 4);
  if (split_ok && (ctx->fft_path == 6 || (!tc && ctx->fft_path == 0 &&
                                          (ctx->matvec_path == 5 || batch >= kTcMinBatch)))) {
    TcPlan sp;
    if (tc_plan_dft(n, d_eff, k_out, &sp)) {
      pl = sp;
      tc = true;
    }
  }
  if (tc) use_matrix = true;
  const void* d_m = nullptr;
  const void* d_tw = nullptr;
  const std::vector<uint32_t>* h_tw = nullptr;
  if (use_matrix) {
    // M[i][j] = omega^(i j), row-major, Montgomery form
    auto gen = [&](std::vector<Fe>& m) {
      Fe w;
      int r = check_omega(ctx, omega, n, w);
      if (r) return r;
      m.assign((size_t)k_out * (d_eff ? d_eff : 1), fe_zero());
      Fe wi = ctx->field->one();  // omega^i
      for (int i = 0; i < k_out; i++) {
        Fe acc = ctx->field->one();
        for (int j = 0; j < d_eff; j++) {
          m[(size_t)i * d_eff + j] = acc;
          acc = ctx->field->mul(acc, wi);
        }
        wi = ctx->field->mul(wi, w);
      }
      return HBG_OK;
    };
    const std::string key = make_key("dft", omega, 32, &n, sizeof n, k_out, d_eff);
    rc = tc && pl.split
             ? get_const(ctx, key + "|tcsplit", &d_m,
                         [&](std::vector<uint32_t>& host) {
                           Fe w;
                           int r = check_omega(ctx, omega, n, w);
                           if (r) return r;
                           tc_build_bmat_dft(*ctx->field, w, d_eff, pl, host);
                           return HBG_OK;
                         })
         : tc ? tc_const(ctx, key, k_out, d_eff, pl, &d_m, gen)
            : get_const(ctx, key, &d_m, [&](std::vector<uint32_t>& host) {
                std::vector<Fe> m;
                int r = gen(m);
                if (r) return r;
                interleave(m, k_out, d_eff, host);
                return HBG_OK;
              });
  } else {
    rc = twiddles(ctx, omega, n, &d_tw, &h_tw);
  }
  if (rc) return rc;
  return run_rows(ctx, polys, (size_t)d * 32, out, (size_t)k_out * 32, batch, mem,
                  [&](const void* di, void* dout, size_t rows) {
                    if (tc)
                      return launch_tc(ctx, d_m, pl, k_out, d_eff, di, (size_t)d * 32, dout,
                                       (size_t)k_out * 32, rows);
                    if (use_matrix) return launch_matvec(ctx, d_m, k_out, d_eff, di, d, dout, k_out, rows);
                    return launch_ntt(ctx, d_tw, h_tw, n, di, d, dout, k_out, rows);
                  });
}

int hbg_fft_batch_interpolate(hbg_ctx* ctx, const uint64_t omega[4], int n, const int32_t* zs, int k,
                              const uint64_t* ys, size_t batch, uint64_t* out, int mem) {
  if (!ctx) return HBG_ERR_INVALID;
  if (!omega || k < 0 || (k && !zs)) return fail(ctx, HBG_ERR_INVALID, "bad size or null points");
  CU(cudaSetDevice(ctx->device));
  { int trc = cache_trim(ctx); if (trc) return trc; }
  // NTT-structured path (fnt_decode_step2): forced, or automatically for k > 128, where the
  // O(k^3) host inverse of the matrix path and its O(k^2) products per row stop paying
  if (k >= 1 && n >= 2 && (ctx->interp_path == 2 || (ctx->interp_path == 0 && k > 128))) {
    FntConst fc;
    int frc = fnt_constants(ctx, omega, n, zs, k, &fc);
    if (frc == HBG_OK) {
      if (batch == 0) return HBG_OK;
      if (!out || !ys) return fail(ctx, HBG_ERR_INVALID, "null batch buffer");
      return run_rows(ctx, ys, (size_t)k * 32, out, (size_t)k * 32, batch, mem,
                      [&](const void* di, void* dout, size_t rows) {
                        return fnt_interpolate(ctx, fc, n, k, di, rows, dout);
                      });
    }
    if (frc != HBG_ERR_UNSUPPORTED || ctx->interp_path == 2) return frc;
  }
  if (k > 4096)
    return fail(ctx, HBG_ERR_UNSUPPORTED, "matrix-path interpolation from more than 4096 points "
                                          "(this field has no root of unity for the NTT path)");
  const void* d_m = nullptr;
  const std::string key = make_key("finv", omega, 32, zs, (size_t)k * 4, n, k);
  int rc = interp_matrix(ctx, key, k, &d_m,
                         [&](std::vector<Fe>& x) {
                           Fe w;
                           int r = check_omega(ctx, omega, n, w);
                           if (r) return r;
                           x.resize(k);
                           for (int i = 0; i < k; i++) {
                             if (zs[i] < 0 || zs[i] >= n)
                               return fail(ctx, HBG_ERR_INVALID, "z outside [0, n)");
                             x[i] = ctx->field->pow_u64(w, (uint64_t)zs[i]);
                           }
                           return HBG_OK;
                         });
  if (rc) return rc;
  if (batch == 0 || k == 0) return HBG_OK;
  if (!out || !ys) return fail(ctx, HBG_ERR_INVALID, "null batch buffer");
  TcPlan pl;
  const bool tc = tc_wanted(ctx, k, k, batch, &pl);
  const void* d_b = nullptr;
  if (tc) {
    rc = tc_const(ctx, key, k, k, pl, &d_b, [&](std::vector<Fe>& inv) {
      Fe w;
      int r = check_omega(ctx, omega, n, w);
      if (r) return r;
      std::vector<Fe> x(k);
      for (int i = 0; i < k; i++) x[i] = ctx->field->pow_u64(w, (uint64_t)zs[i]);
      return vandermonde_inverse(*ctx->field, x, inv) ? HBG_OK : HBG_ERR_SINGULAR;
    });
    if (rc) return rc;
  }
  return run_rows(ctx, ys, (size_t)k * 32, out, (size_t)k * 32, batch, mem,
                  [&](const void* di, void* dout, size_t rows) {
                    if (tc) return launch_tc(ctx, d_b, pl, k, k, di, (size_t)k * 32, dout, (size_t)k * 32, rows);
                    return launch_interp(ctx, key, d_m, k, di, dout, rows);
                  });
}

int hbg_ctx_set_cache_limit(hbg_ctx* ctx, size_t bytes) {
  if (!ctx) return HBG_ERR_INVALID;
  ctx->cache_limit = bytes;
  return HBG_OK;
}

int hbg_ctx_set_sm_limit(hbg_ctx* ctx, int ctas) {
  if (!ctx) return HBG_ERR_INVALID;
  if (ctas < 0) return fail(ctx, HBG_ERR_INVALID, "sm limit must be >= 0");
  ctx->sm_limit = ctas;
  return HBG_OK;
}

int hbg_columns_to_rows(hbg_ctx* ctx, const uint64_t* colbuf, size_t batch, const int32_t* idx, int k,
                        uint64_t* rows) {
  if (!ctx) return HBG_ERR_INVALID;
  if (k < 0 || k > 256 || (k && !idx)) return fail(ctx, HBG_ERR_INVALID, "0 <= k <= 256 columns");
  if (batch == 0 || k == 0) return HBG_OK;
  if (!colbuf || !rows) return fail(ctx, HBG_ERR_INVALID, "null batch buffer");
  CU(cudaSetDevice(ctx->device));
  ColumnIdx ci;
  memset(&ci, 0, sizeof ci);
  for (int j = 0; j < k; j++) {
    if (idx[j] < 0) return fail(ctx, HBG_ERR_INVALID, "negative column index");
    ci.idx[j] = idx[j];
  }
  unsigned long long blocks = (batch * (unsigned long long)k + 255) / 256;
  const unsigned long long cap = (unsigned long long)ctx->sm_count * 8;
  if (blocks > cap) blocks = cap;
  columns_to_rows_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>((const uint4*)colbuf, batch, ci, k,
                                                                     (uint4*)rows);
  CU(cudaGetLastError());
  ctx->launches++;
  ctx->last_kernel = "columns_to_rows_kernel";
  return HBG_OK;
}

int hbg_compare_columns(hbg_ctx* ctx, const uint64_t* rows, int row_width, int col_offset,
                        const uint64_t* colbuf, size_t batch, const int32_t* idx, int m, int32_t* flags_dev,
                        int32_t* flags_host) {
  if (!ctx) return HBG_ERR_INVALID;
  if (m < 0 || m > 256 || (m && !idx) || row_width < 1 || col_offset < 0)
    return fail(ctx, HBG_ERR_INVALID, "bad size (0 <= m <= 256 columns)");
  if (m == 0) return HBG_OK;
  if (!flags_dev) return fail(ctx, HBG_ERR_INVALID, "null flags buffer");
  CU(cudaSetDevice(ctx->device));
  CU(cudaMemsetAsync(flags_dev, 0, (size_t)m * 4, ctx->stream));
  if (batch) {
    if (!rows || !colbuf) return fail(ctx, HBG_ERR_INVALID, "null batch buffer");
    ColumnIdx ci;
    memset(&ci, 0, sizeof ci);
    for (int j = 0; j < m; j++) {
      if (idx[j] < 0 || col_offset + idx[j] >= row_width)
        return fail(ctx, HBG_ERR_INVALID, "column index outside the row");
      ci.idx[j] = idx[j];
    }
    unsigned long long blocks = (batch * (unsigned long long)m + 255) / 256;
    const unsigned long long cap = (unsigned long long)ctx->sm_count * 8;
    if (blocks > cap) blocks = cap;
    compare_columns_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(
        (const uint4*)rows, row_width, col_offset, (const uint4*)colbuf, batch, ci, m, flags_dev);
    CU(cudaGetLastError());
    ctx->launches++;
    ctx->last_kernel = "compare_columns_kernel";
  }
  if (flags_host) {
    CU(cudaMemcpyAsync(flags_host, flags_dev, (size_t)m * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
  }
  return HBG_OK;
}

int hbg_interpolate_reencode(hbg_ctx* ctx, const uint64_t* xs_k, int k, const uint64_t* xs_all, int n,
                             const uint64_t* ys, size_t batch, uint64_t* out, int mem) {
  if (!ctx) return HBG_ERR_INVALID;
  if (k < 1 || n < 0 || !xs_k || (n && !xs_all)) return fail(ctx, HBG_ERR_INVALID, "bad size or null points");
  CU(cudaSetDevice(ctx->device));
  { int trc = cache_trim(ctx); if (trc) return trc; }
  // the stacked matrix [W ; V(xs_all) W], W = V(xs_k)^-1, row-major, Montgomery form
  auto gen = [&](std::vector<Fe>& m) {
    std::vector<Fe> xk, xa, inv;
    int r = load_points(ctx, xs_k, k, xk);
    if (r) return r;
    r = load_points(ctx, xs_all, n, xa);
    if (r) return r;
    if (!vandermonde_inverse(*ctx->field, xk, inv))
      return fail(ctx, HBG_ERR_SINGULAR, "evaluation points are not pairwise distinct");
    const HostField& f = *ctx->field;
    m.assign((size_t)(k + n) * k, fe_zero());
    for (int i = 0; i < k * k; i++) m[i] = inv[i];
    for (int i = 0; i < n; i++) {  // row i of V(xs_all) W by Horner over the rows of W
      for (int j = 0; j < k; j++) {
        Fe acc = inv[(size_t)(k - 1) * k + j];
        for (int l = k - 2; l >= 0; l--) acc = f.add(f.mul(acc, xa[i]), inv[(size_t)l * k + j]);
        m[(size_t)(k + i) * k + j] = acc;
      }
    }
    return HBG_OK;
  };
  const int n_out = k + n;
  std::string key = make_key("ireenc", xs_k, (size_t)k * 32, xs_all, (size_t)n * 32, k, n);
  TcPlan pl;
  const bool tc = tc_wanted(ctx, n_out, k, batch, &pl);
  const void* d_m = nullptr;
  int rc = tc ? tc_const(ctx, key, n_out, k, pl, &d_m, gen)
              : get_const(ctx, key, &d_m, [&](std::vector<uint32_t>& host) {
                  std::vector<Fe> m;
                  int r = gen(m);
                  if (r) return r;
                  interleave(m, n_out, k, host);
                  return HBG_OK;
                });
  if (rc) return rc;
  if (batch == 0) return HBG_OK;
  if (!ys || !out) return fail(ctx, HBG_ERR_INVALID, "null batch buffer");
  return run_rows(ctx, ys, (size_t)k * 32, out, (size_t)n_out * 32, batch, mem,
                  [&](const void* di, void* dout, size_t rows) {
                    if (tc)
                      return launch_tc(ctx, d_m, pl, n_out, k, di, (size_t)k * 32, dout, (size_t)n_out * 32,
                                       rows);
                    return launch_matvec(ctx, d_m, n_out, k, di, k, dout, n_out, rows);
                  });
}

}  // extern "C"
