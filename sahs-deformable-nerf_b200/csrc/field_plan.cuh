// Network plan of the fused field kernel: how the reference's three MLPs (WarpFieldMLP, HyperSheetMLP,
// NeRFMLP; ref: nerf/modules.py:168-295, :323-390, :401-462) are cut into tensor-core "stages".
//
// Data layout
//   * activation buffer X (shared memory): 4 K-chunks of [128 points x 64 features] 16-bit (fp16), each chunk a UMMA
//     K-major SWIZZLE_128B operand tile (16 KB).  Chunk c holds features 64c..64c+63.
//   * a stage = one B operand block [n rows (outputs) x 64 inputs] fp16 in the same swizzled layout
//     (n*128 bytes), streamed by the TMA unit from the packed weight image in consumption order.
//   * accumulators: TMEM columns 0..255 (fp32), row i of the tile <-> TMEM lane i.
//   * a pass = the stages issued between two worker phases (epilogue / encoding writes).
#pragma once
#include <stdint.h>
#include "../../include/sahs_b200.h"

constexpr int kTileRows = 128;
constexpr int kChunkCols = 64;
constexpr int kChunkBytes = kTileRows * kChunkCols * 2;  // 16 KB
constexpr int kMaxStages = 160;
constexpr int kStageSlotBytes = 16384;

// Operand format of trunk and heads on the RENDER path: fp16 by default (11-bit significand, 8x finer than bf16 at the
// same MMA rate; conversions saturate at +-65504), bf16 when the library is built with -DSAHS_RENDER_BF16 (`make BF16=1`:
// the format the task statement names; range like fp32, 8-bit significand).  The deformation phase stays fp16 (its output
// feeds a 2^(L-1)-amplifying encoding) and so does training (tapes, dgrad, wgrad), which scales its gradients into range.
#ifdef SAHS_RENDER_BF16
constexpr bool kRenderTrunkF16 = false;
#else
constexpr bool kRenderTrunkF16 = true;
#endif
enum : uint8_t { ST_WAIT_A = 1, ST_COMMIT = 2, ST_FRESH = 4, ST_F16 = 8, ST_WIDE = 16 };  // F16: fp16 operands (else bf16)
// ST_WIDE: the B block is [256 outputs x 32 inputs] (K-major rows of 64 bytes, SWIZZLE_64B, 16 KB) instead of
// [n <= 128 outputs x 64 inputs] (128-byte rows, SWIZZLE_128B): one N = 256 MMA per K = 16 step.  An N = 128 MMA reads
// 4 KB of A + 4 KB of B per 64 tensor cycles -- with the half of B it serves to the peer CTA that is 117 B/cycle of the
// SM's 128 B/cycle of shared-memory bandwidth, so every TMA write and epilogue store slows the MMAs (measured 92 cycles
// per MMA in the kernel against 70 alone); N = 256 reads A once for twice the work: 96 B/cycle.

struct StageRec {   // 16 bytes, lives in kernel parameter space
  uint8_t n8;       // N / 8
  uint8_t kflags;   // ksteps (low 3 bits) | flags << 3
  uint8_t a_chunk;  // which X chunk is the A operand
  uint8_t d_col8;   // accumulator column offset / 8
  uint8_t a_chunk2; // second A chunk multiplied by the same stage (split precision: the lo part), 0xFF = none
  uint8_t a_k16;    // first K = 16 step of the A chunk this stage multiplies (ST_WIDE stages cover half a chunk: 0 or 2)
  uint8_t pad[2];
  // derived on the host (sahs_finalize_plan) so that the MMA issue loop decodes a stage with one 16-byte load:
  uint32_t idesc;   // tcgen05 instruction descriptor for M = 128 (the pair kernel ORs in M = 256)
  uint16_t a_off;   // a_chunk  * kChunkBytes / 16: offset of the A descriptor's address field
  uint16_t a2_off;  // a_chunk2 * kChunkBytes / 16 (unused when a_chunk2 == 0xFF)
};

// kind::f16 instruction descriptor (D fp32, A/B both fp16 (format 0) or bf16 (1), K-major): N at bit 17, M at bit 24
inline uint32_t sahs_idesc_m128(uint32_t n, bool f16) {
  const uint32_t fmt = f16 ? 0u : 1u;
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}

struct alignas(8) FieldPlan {   // (8: the issue loops read a StageRec as two 8-byte constant loads)
  int32_t num_stages;
  int32_t total_bytes;  // packed image bytes of one level
  StageRec st[kMaxStages];
};

struct PackSrc {      // part of a stage image copied from one fp32 [out,in] weight matrix
  const float* w;     // device pointer (NULL: unused)
  int32_t ld;         // in_features
  int32_t src_row0, src_col0;
  int32_t dst_row0, nrows, ncols;
  int32_t dst_col0;   // first image column the block lands in (0 except for the fused output layer of the dgrad plan)
  int32_t transpose;  // image(row, col) = w[(src_row0 + col) * ld + src_col0 + (row - dst_row0)]  (dgrad images)
};
struct PackStage {
  PackSrc src[2];
  uint32_t dst_off;   // byte offset in the packed image
  int32_t n;          // rows of the image (zero padded)
  int32_t f16;        // store fp16 instead of bf16
  int32_t lo;         // store the residual fp16(w - fp16(w)) (split-precision deformation phase)
  int32_t wide;       // ST_WIDE image: rows of 32 inputs (64 bytes), SWIZZLE_64B
};

// fp32 per-frame constant block: offsets in floats
struct FoldSection {
  const float* bias;  // [n]
  const float* w;     // [n, ld] weight whose constant columns are folded (NULL: plain copy of bias)
  int32_t ld, col0, ncols, c_off;  // fold W[:, col0:col0+ncols] @ cvec[c_off:c_off+ncols]
  int32_t n, dst;
};
struct CopySection {  // small fp32 head weights copied verbatim
  const float* src;
  int32_t count, dst;
};

// Dimensions derived from the spec, shared by host plan builder and device worker code.
struct NetDims {
  int e0_dim, e0_k, e0_chunks;      // PE(xyz): 63->64 | 93->96
  int amb_pe;                       // ambient PE width
  int e1_dim, e1_k, e1_chunks;      // PE(mapped xyz) | PE(ambient)
  int wh, hh, whh;                  // warp hidden, hyper hidden, merged width
  int w_layers, w_skip;             // merged deformation net depth / skip layer (warp & hyper agree)
  int e0_resident;                  // E0 stays in X next to the hidden activations during the W phase
  int e0_chunk_base;                // first X chunk of E0 for layer 0 / skip pass
  int th, t_layers, t_skip;
  int hd;                           // head hidden (th/2)
  int xtra_dim, xtra_k;             // dir PE (+ grid) appended to feat for layers_dir.0
  int ct_off, ct_len;               // trunk constant vector = cvec[ct_off : ct_off+ct_len]
  int use_w;                        // deformation phase present
  int w_split;                      // deformation phase in split precision (fp16 hi + lo), warp and hyper nets run
                                    // one after the other: needed when the encoding has more than 10 octaves
  // frame-constant block offsets (floats)
  int off_wbias, off_wfinal, off_tbias, off_featb, off_alpha, off_hbias, off_outb, fc_total;
  // training tapes (fp16 chunk images, see sahs_make_dims): column offsets of the saved activations (tx_*) and of the
  // activation gradients (td_*), all multiples of 64; sign masks are indexed by layer (W i -> i, T i -> w_layers + i, H i -> .. + t_layers + i)
  int tx_e0, tx_wh, tx_e1, tx_th, tx_feat, tx_xtra, tx_hh, tx_total;
  int td_wh, td_final, td_th, td_feat, td_hh, td_out, td_total;
  int n_mask_layers;
};

// bytes of a stage's B image (both CTAs' halves together)
#ifdef __CUDACC__
__host__ __device__
#endif
inline uint32_t sahs_stage_bytes(const StageRec& r) {
  return (uint32_t)r.n8 * (((r.kflags >> 3) & ST_WIDE) ? 512u : 1024u);
}

inline int sahs_round_up(int v, int m) { return (v + m - 1) / m * m; }

// returns 0 on success
inline int sahs_make_dims(const sahs_model_spec& s, NetDims& d, bool train = false) {
  d = NetDims{};
  d.e0_dim = (s.xyz_inc ? 3 : 0) + 6 * s.xyz_L;
  d.e0_k = sahs_round_up(d.e0_dim, 16);
  d.e0_chunks = (d.e0_k + 63) / 64;
  d.amb_pe = s.use_ambient ? ((s.amb_inc ? s.amb_dim : 0) + 2 * s.amb_dim * s.amb_L) : 0;
  d.e1_dim = d.e0_dim + d.amb_pe;
  d.e1_k = sahs_round_up(d.e1_dim, 16);
  d.e1_chunks = (d.e1_k + 63) / 64;
  d.use_w = (s.use_warp && s.use_ambient) ? 1 : 0;
  if (s.use_warp != s.use_ambient) return -1;  // shipped configs enable both or neither
  d.wh = s.warp_hidden;
  d.hh = s.hyper_hidden;
  d.whh = d.wh + d.hh;
  d.w_layers = s.warp_layers;
  d.w_skip = s.warp_skip;
  if (d.use_w) {
    if (s.hyper_layers != s.warp_layers || s.hyper_skip != s.warp_skip) return -2;
    if (d.wh % 64 || d.hh % 64 || d.wh > 128 || d.hh > 128 || d.whh > 256) return -3;
    if (d.w_skip <= 0 || d.w_skip >= d.w_layers) return -4;
  }
  d.w_split = (d.use_w && s.xyz_L > 10 && !train) ? 1 : 0;   // training always uses the merged fp16 phase
  if (d.w_split && (d.wh != 128 || d.hh != 64)) return -10;
  d.e0_resident = (!d.w_split && d.whh / 64 + d.e0_chunks <= 4) ? 1 : 0;
  d.e0_chunk_base = d.e0_resident ? d.whh / 64 : 0;
  d.th = s.trunk_hidden;
  d.t_layers = s.trunk_layers;
  d.t_skip = s.trunk_skip;
  if (d.th != 256) return -5;
  if (d.t_skip <= 0 || d.t_skip >= d.t_layers) return -6;
  if (d.e0_chunks > 2 || d.e1_chunks > 2) return -7;
  d.hd = d.th / 2;
  d.xtra_dim = (s.dir_inc ? 3 : 0) + 6 * s.dir_L + (s.use_grid ? SAHS_GRID_CH : 0);
  d.xtra_k = sahs_round_up(d.xtra_dim, 16);
  if (d.xtra_k > 64) return -8;
  d.ct_off = s.trunk_driving ? 0 : SAHS_DRIVING_DIM;
  d.ct_len = (s.trunk_driving ? SAHS_DRIVING_DIM : 0) + (s.trunk_pose ? SAHS_POSE_CODE_DIM : 0);
  if (s.amb_dim > 4 || s.amb_dim < 0) return -9;
  int o = 0;
  d.off_wbias = o;  o += d.use_w ? d.w_layers * d.whh : 0;
  d.off_wfinal = o; o += d.use_w ? (3 * d.wh + 4 + s.amb_dim * d.hh + 4) : 0;   // [wf | bf(4) | wa | ba(4)]
  d.off_tbias = o;  o += d.t_layers * d.th;
  d.off_featb = o;  o += d.th;
  d.off_alpha = o;  o += d.th + 4;
  d.off_hbias = o;  o += 4 * 2 * d.hd;
  d.off_outb = o;   o += 16;
  d.fc_total = o;
  // Training tapes are stored as tile-major chunk images: tape[tile][slot][128 points x 64 columns, 128B-swizzled] --
  // exactly the shared-memory operand chunks, written by TMA bulk stores and read back by the wgrad kernel with bulk
  // loads.  Every item therefore starts at a multiple of 64 columns (slot = column / 64).
  auto up64 = [](int v) { return (v + 63) / 64 * 64; };
  int t = 0;
  d.tx_e0 = t;   t += d.use_w ? up64(d.e0_k) : 0;
  d.tx_wh = t;   t += d.use_w ? d.w_layers * up64(d.whh) : 0;
  d.tx_e1 = t;   t += up64(d.e1_k);
  d.tx_th = t;   t += d.t_layers * d.th;
  d.tx_feat = t; t += d.th;
  d.tx_xtra = t; t += 64;
  d.tx_hh = t;   t += 4 * 2 * d.hd;
  d.tx_total = t;
  t = 0;
  d.td_wh = t;    t += d.use_w ? d.w_layers * up64(d.whh) : 0;
  d.td_final = t; t += d.use_w ? 64 : 0;
  d.td_th = t;    t += d.t_layers * d.th;
  d.td_feat = t;  t += d.th;
  d.td_hh = t;    t += 4 * 2 * d.hd;
  d.td_out = t;   t += 64;
  t += 64;        // the wgrad kernel always fetches two dY chunks: one slot of slack after the last item
  d.td_total = t;
  d.n_mask_layers = (d.use_w ? d.w_layers : 0) + d.t_layers + 4;
  return 0;
}

// host-only: plan + pack/fold descriptors (field_host.cu)
struct HostPlan {
  NetDims dims;
  FieldPlan plan;
  PackStage pack[kMaxStages];
  FoldSection fold[64];
  int num_fold;
  CopySection copy[8];
  int num_copy;
};
int sahs_build_host_plan(const sahs_model_spec& spec, const float* const* params, HostPlan& hp, bool train = false);
// dgrad plan: stages hold transposed fp16 weight blocks in the order the backward kernel consumes them
int sahs_build_bwd_plan(const sahs_model_spec& spec, const float* const* params, HostPlan& hp);
