// Host side of the fused field path: plan construction, weight packing (fp32 state_dict -> 16-bit (fp16; bf16 when a stage is not flagged f16) swizzled stage
// images), per-frame constant folding.  See field_plan.cuh for the layout.
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <cuda_fp16.h>
#include "sahs_common.cuh"
#include "field_plan.cuh"

namespace {

struct ParamIndex {
  int grid = 0;
  int warp_w[16], warp_b[16], warp_fw = -1, warp_fb = -1;
  int hyp_w[16], hyp_b[16], hyp_fw = -1, hyp_fb = -1;
  int trunk_w[16], trunk_b[16];
  int feat_w, feat_b, alpha_w, alpha_b;
  int dir_w[4], dir_b[4], rgb_w, rgb_b;
  int seg_w[4], seg_b[4], segf_w, segf_b;
  int count = 0;
};

bool index_params(const sahs_model_spec& s, ParamIndex& pi) {
  if (s.warp_layers > 16 || s.hyper_layers > 16 || s.trunk_layers > 16) return false;
  int k = 0;
  pi.grid = k++;
  if (s.use_warp) {
    for (int i = 0; i < s.warp_layers; ++i) { pi.warp_w[i] = k++; pi.warp_b[i] = k++; }
    pi.warp_fw = k++; pi.warp_fb = k++;
  }
  if (s.use_ambient) {
    for (int i = 0; i < s.hyper_layers; ++i) { pi.hyp_w[i] = k++; pi.hyp_b[i] = k++; }
    pi.hyp_fw = k++; pi.hyp_fb = k++;
  }
  for (int i = 0; i < s.trunk_layers; ++i) { pi.trunk_w[i] = k++; pi.trunk_b[i] = k++; }
  pi.feat_w = k++; pi.feat_b = k++; pi.alpha_w = k++; pi.alpha_b = k++;
  for (int i = 0; i < 4; ++i) { pi.dir_w[i] = k++; pi.dir_b[i] = k++; }
  pi.rgb_w = k++; pi.rgb_b = k++;
  for (int i = 0; i < 4; ++i) { pi.seg_w[i] = k++; pi.seg_b[i] = k++; }
  pi.segf_w = k++; pi.segf_b = k++;
  pi.count = k;
  return true;
}

struct Builder {
  HostPlan& hp;
  const float* const* P;
  int ns = 0;
  uint32_t off = 0;
  bool pass_open = false;
  bool f16 = false;   // operand format of the stages being emitted
  bool lo = false;    // emit the residual image fp16(w - fp16(w))
  int a_chunk2 = 0xFF;
  bool wide = false;  // emit ST_WIDE stages where a layer is 256 outputs wide (render plans; see field_plan.cuh)

  // one stage: rows [row0,row0+n) x cols [col0, col0+kvalid) of weight `pidx` (ld = in_features)
  void stage(int pidx, int ld, int row0, int n, int col0, int kvalid, int a_chunk, int d_col, bool fresh) {
    StageRec& r = hp.plan.st[ns];
    int ksteps = (kvalid + 15) / 16;
    r.n8 = (uint8_t)(n / 8);
    uint8_t flags = (fresh ? ST_FRESH : 0) | (pass_open ? 0 : ST_WAIT_A) | (f16 ? ST_F16 : 0);
    r.kflags = (uint8_t)(ksteps | (flags << 3));
    r.a_chunk = (uint8_t)a_chunk;
    r.d_col8 = (uint8_t)(d_col / 8);
    r.a_chunk2 = (uint8_t)a_chunk2;
    r.a_k16 = 0;
    r.pad[0] = r.pad[1] = 0;
    PackStage& ps = hp.pack[ns];
    memset(&ps, 0, sizeof(ps));
    ps.dst_off = off;
    ps.n = n;
    ps.f16 = f16 ? 1 : 0;
    ps.lo = lo ? 1 : 0;
    ps.src[0].w = P ? P[pidx] : nullptr;
    ps.src[0].ld = ld;
    ps.src[0].src_row0 = row0; ps.src[0].src_col0 = col0;
    ps.src[0].dst_row0 = 0; ps.src[0].nrows = n; ps.src[0].ncols = kvalid;
    off += (uint32_t)n * 128u;
    pass_open = true;
    ++ns;
  }
  // ST_WIDE stage: all 256 outputs x inputs [col0, col0 + kvalid) of weight `pidx`, kvalid <= 32; the A operand is the
  // half `khalf` of chunk a_chunk
  void wstage(int pidx, int ld, int col0, int kvalid, int a_chunk, int khalf, bool fresh) {
    const uint32_t off0 = off;
    stage(pidx, ld, 0, 256, col0, kvalid, a_chunk, 0, fresh);
    StageRec& r = hp.plan.st[ns - 1];
    r.kflags |= (uint8_t)(ST_WIDE << 3);
    r.a_k16 = (uint8_t)(2 * khalf);
    hp.pack[ns - 1].wide = 1;
    off = off0 + 256u * 64u;
  }
  // a 256-output layer block over inputs [col0, col0 + kdim) read from X chunks a_chunk0..: wide stages when enabled,
  // else the two 128-row halves one after the other (same K order per output either way: bit-identical results)
  void block256(int pidx, int ld, int col0, int kdim, int a_chunk0, bool fresh_first) {
    const int nch = (kdim + 63) / 64;
    if (wide) {
      bool fresh = fresh_first;
      for (int kc = 0; kc < nch; ++kc)
        for (int h = 0; h < 2; ++h) {
          int kv = kdim - 64 * kc - 32 * h;
          if (kv <= 0) continue;
          if (kv > 32) kv = 32;
          wstage(pidx, ld, col0 + 64 * kc + 32 * h, kv, a_chunk0 + kc, h, fresh);
          fresh = false;
        }
      return;
    }
    for (int half = 0; half < 2; ++half)
      for (int kc = 0; kc < nch; ++kc) {
        int kv = kdim - 64 * kc; if (kv > 64) kv = 64;
        stage(pidx, ld, 128 * half, 128, col0 + 64 * kc, kv, a_chunk0 + kc, 128 * half, fresh_first && kc == 0);
      }
  }
  // dgrad stage: image rows = input features [in_col0, in_col0 + n_valid) (zero padded to n), image columns
  // [dst_col0, dst_col0 + k_valid) = output features [out_row0, out_row0 + k_valid); the A operand is the 64-wide chunk
  // of output-feature gradients `a_chunk`.
  void tstage(int pidx, int ld, int out_row0, int k_valid, int dst_col0, int in_col0, int n, int n_valid, int a_chunk,
              int d_col, bool fresh) {
    stage(pidx, ld, 0, n, 0, dst_col0 + k_valid, a_chunk, d_col, fresh);
    PackSrc& src = hp.pack[ns - 1].src[0];
    src.transpose = 1;
    src.src_row0 = out_row0; src.src_col0 = in_col0;
    src.dst_row0 = 0; src.nrows = n_valid; src.ncols = k_valid; src.dst_col0 = dst_col0;
  }
  void end_pass() {
    StageRec& r = hp.plan.st[ns - 1];
    r.kflags |= (uint8_t)(ST_COMMIT << 3);
    pass_open = false;
  }
};

// fills the derived StageRec fields the MMA issue loop reads
void finalize_plan(FieldPlan& plan) {
  for (int i = 0; i < plan.num_stages; ++i) {
    StageRec& r = plan.st[i];
    const uint32_t flags = r.kflags >> 3;
    r.idesc = sahs_idesc_m128((uint32_t)r.n8 * 8u, (flags & ST_F16) != 0);
    r.a_off = (uint16_t)(r.a_chunk * (kChunkBytes >> 4) + r.a_k16 * 2);   // 16-byte units; a K = 16 step is 32 bytes
    r.a2_off = (uint16_t)((r.a_chunk2 == 0xFF ? 0 : r.a_chunk2) * (kChunkBytes >> 4));
  }
}

}  // namespace

extern "C" int sahs_param_count(const sahs_model_spec* spec) {
  ParamIndex pi;
  if (!spec || !index_params(*spec, pi)) return SAHS_EINVAL;
  return pi.count;
}

static int build_host_plan_uncached(const sahs_model_spec& s, const float* const* params, HostPlan& hp, bool train);

// The launch paths call this once per kernel launch with params == NULL (they need the stage table and the dimensions,
// not the pack sources): that result depends on the spec alone, so it is built once per (thread, spec, train) and copied.
int sahs_build_host_plan(const sahs_model_spec& s, const float* const* params, HostPlan& hp, bool train) {
  if (params) return build_host_plan_uncached(s, params, hp, train);
  struct Cached { bool valid = false; sahs_model_spec spec; HostPlan hp; };
  static thread_local Cached cache[2];
  Cached& c = cache[train ? 1 : 0];
  if (!c.valid || memcmp(&c.spec, &s, sizeof(s)) != 0) {
    c.valid = false;
    int rc = build_host_plan_uncached(s, nullptr, c.hp, train);
    if (rc) return rc;
    c.spec = s;
    c.valid = true;
  }
  if (&hp != &c.hp) hp = c.hp;
  return SAHS_OK;
}

static int build_host_plan_uncached(const sahs_model_spec& s, const float* const* params, HostPlan& hp, bool train) {
  NetDims& d = hp.dims;
  int rc = sahs_make_dims(s, d, train);
  if (rc) {
    sahs_set_error("unsupported model spec (dims check %d)", rc);
    return SAHS_EUNSUPPORTED;
  }
  ParamIndex pi;
  if (!index_params(s, pi)) return SAHS_EINVAL;
  Builder b{hp, params};
  b.f16 = true;   // all operands fp16 (see field_fwd.cu "Precision")
  // (training keeps the narrow stages; SAHS_FIELD_WIDE=0, read once per process: A/B measurements)
  static const bool wide_on = [] { const char* e = getenv("SAHS_FIELD_WIDE"); return !(e && e[0] == '0'); }();
  b.wide = wide_on && !train && d.th == 256;
  hp.num_fold = 0;
  hp.num_copy = 0;
  auto P = [&](int i) -> const float* { return params ? params[i] : nullptr; };
  auto fold = [&](int bias_i, int w_i, int ld, int col0, int ncols, int c_off, int n, int dst) {
    FoldSection& f = hp.fold[hp.num_fold++];
    f.bias = P(bias_i); f.w = (w_i >= 0 && ncols > 0) ? P(w_i) : nullptr;
    f.ld = ld; f.col0 = col0; f.ncols = ncols; f.c_off = c_off; f.n = n; f.dst = dst;
  };
  const int CW = SAHS_DRIVING_DIM + SAHS_POSE_CODE_DIM;  // 112 frame-constant inputs of warp/hyper
  // ---------------- deformation phase: warp | hyper merged -------------------------------------------
  if (d.use_w && d.w_split) {
    // split precision: activations and weights as fp16 hi + lo, products hi*hi + lo*hi + hi*lo (fp32 accumulate).
    // X chunks: hidden hi at [0, n/64), lo at [n/64, 2n/64); encoding hi at [0,2), lo at [2,4).
    b.f16 = true;
    const int in0 = d.e0_dim + CW;
    for (int net = 0; net < 2; ++net) {
      const int n = net == 0 ? d.wh : d.hh;
      const int* wi = net == 0 ? pi.warp_w : pi.hyp_w;
      const int* bi = net == 0 ? pi.warp_b : pi.hyp_b;
      const int boff = net == 0 ? 0 : d.wh;
      auto part = [&](int pidx, int ld, int col0, int kvalid, int hi0, int lo0, bool& fresh) {
        for (int kc = 0; kc * 64 < kvalid; ++kc) {
          int kv = kvalid - 64 * kc; if (kv > 64) kv = 64;
          b.lo = false; b.a_chunk2 = lo0 + kc;
          b.stage(pidx, ld, 0, n, col0 + 64 * kc, kv, hi0 + kc, 0, fresh); fresh = false;
          b.lo = true; b.a_chunk2 = 0xFF;
          b.stage(pidx, ld, 0, n, col0 + 64 * kc, kv, hi0 + kc, 0, false);
        }
        b.lo = false; b.a_chunk2 = 0xFF;
      };
      for (int i = 0; i < d.w_layers; ++i) {
        const bool first = i == 0, skip = i == d.w_skip;
        const int ld = first ? in0 : (skip ? n + in0 : n);
        bool fresh = true;
        if (!first) part(wi[i], ld, 0, n, 0, n / 64, fresh);
        if (first || skip) {
          if (skip) b.end_pass();
          part(wi[i], ld, first ? 0 : n, d.e0_dim, 0, 2, fresh);
          fold(bi[i], wi[i], ld, (first ? 0 : n) + d.e0_dim, CW, 0, n, d.off_wbias + i * d.whh + boff);
        } else {
          fold(bi[i], -1, 0, 0, 0, 0, n, d.off_wbias + i * d.whh + boff);
        }
        b.end_pass();
      }
    }
    int o = d.off_wfinal;
    hp.copy[hp.num_copy++] = CopySection{P(pi.warp_fw), 3 * d.wh, o}; o += 3 * d.wh;
    hp.copy[hp.num_copy++] = CopySection{P(pi.warp_fb), 3, o}; o += 4;
    hp.copy[hp.num_copy++] = CopySection{P(pi.hyp_fw), s.amb_dim * d.hh, o}; o += s.amb_dim * d.hh;
    hp.copy[hp.num_copy++] = CopySection{P(pi.hyp_fb), s.amb_dim, o};
  } else if (d.use_w) {
    b.f16 = true;   // fp16 operands: the encoding of the warped point amplifies coordinate error by 2^(L-1)
    const int in0 = d.e0_dim + CW;
    for (int i = 0; i < d.w_layers; ++i) {
      const bool first = i == 0, skip = i == d.w_skip;
      const int ldw = first ? in0 : (skip ? d.wh + in0 : d.wh);
      const int ldh = first ? in0 : (skip ? d.hh + in0 : d.hh);
      bool fresh_w = true, fresh_h = true;
      if (!first) {
        for (int kc = 0; kc < d.wh / 64; ++kc) { b.stage(pi.warp_w[i], ldw, 0, d.wh, 64 * kc, 64, kc, 0, fresh_w); fresh_w = false; }
        for (int kc = 0; kc < d.hh / 64; ++kc) { b.stage(pi.hyp_w[i], ldh, 0, d.hh, 64 * kc, 64, d.wh / 64 + kc, d.wh, fresh_h); fresh_h = false; }
      }
      if (first || skip) {
        if (skip && !d.e0_resident) b.end_pass();  // x pass first, encoding pass after the workers rewrote X
        const int cw0 = first ? 0 : d.wh, ch0 = first ? 0 : d.hh;
        for (int kc = 0; kc < d.e0_chunks; ++kc) {
          int kv = d.e0_dim - 64 * kc; if (kv > 64) kv = 64;
          b.stage(pi.warp_w[i], ldw, 0, d.wh, cw0 + 64 * kc, kv, d.e0_chunk_base + kc, 0, fresh_w); fresh_w = false;
        }
        for (int kc = 0; kc < d.e0_chunks; ++kc) {
          int kv = d.e0_dim - 64 * kc; if (kv > 64) kv = 64;
          b.stage(pi.hyp_w[i], ldh, 0, d.hh, ch0 + 64 * kc, kv, d.e0_chunk_base + kc, d.wh, fresh_h); fresh_h = false;
        }
        fold(pi.warp_b[i], pi.warp_w[i], ldw, (first ? 0 : d.wh) + d.e0_dim, CW, 0, d.wh, d.off_wbias + i * d.whh);
        fold(pi.hyp_b[i], pi.hyp_w[i], ldh, (first ? 0 : d.hh) + d.e0_dim, CW, 0, d.hh, d.off_wbias + i * d.whh + d.wh);
      } else {
        fold(pi.warp_b[i], -1, 0, 0, 0, 0, d.wh, d.off_wbias + i * d.whh);
        fold(pi.hyp_b[i], -1, 0, 0, 0, 0, d.hh, d.off_wbias + i * d.whh + d.wh);
      }
      b.end_pass();
    }
    int o = d.off_wfinal;
    hp.copy[hp.num_copy++] = CopySection{P(pi.warp_fw), 3 * d.wh, o}; o += 3 * d.wh;
    hp.copy[hp.num_copy++] = CopySection{P(pi.warp_fb), 3, o}; o += 4;
    hp.copy[hp.num_copy++] = CopySection{P(pi.hyp_fw), s.amb_dim * d.hh, o}; o += s.amb_dim * d.hh;
    hp.copy[hp.num_copy++] = CopySection{P(pi.hyp_fb), s.amb_dim, o};
  }
  // ---------------- trunk ---------------------------------------------------------------------------
  b.f16 = train || kRenderTrunkF16;   // trunk and heads: fp16, or bf16 on the render path of a -DSAHS_RENDER_BF16 build
  {
    const int tin = d.e1_dim + d.ct_len;
    for (int i = 0; i < d.t_layers; ++i) {
      const bool first = i == 0, skip = i == d.t_skip;
      const int ld = first ? tin : (skip ? d.th + tin : d.th);
      if (!first) b.block256(pi.trunk_w[i], ld, 0, d.th, 0, true);
      if (first || skip) {
        if (skip) b.end_pass();
        const int c0 = first ? 0 : d.th;
        b.block256(pi.trunk_w[i], ld, c0, d.e1_dim, 0, first);
        fold(pi.trunk_b[i], pi.trunk_w[i], ld, c0 + d.e1_dim, d.ct_len, d.ct_off, d.th, d.off_tbias + i * d.th);
      } else {
        fold(pi.trunk_b[i], -1, 0, 0, 0, 0, d.th, d.off_tbias + i * d.th);
      }
      b.end_pass();
    }
    b.block256(pi.feat_w, d.th, 0, d.th, 0, true);
    b.end_pass();
    fold(pi.feat_b, -1, 0, 0, 0, 0, d.th, d.off_featb);
    hp.copy[hp.num_copy++] = CopySection{P(pi.alpha_w), d.th, d.off_alpha};
    hp.copy[hp.num_copy++] = CopySection{P(pi.alpha_b), 1, d.off_alpha + d.th};
  }
  // ---------------- heads ---------------------------------------------------------------------------
  {
    const int ld0 = d.th + d.xtra_dim;
    for (int kc = 0; kc < d.th / 64; ++kc) b.stage(pi.dir_w[0], ld0, 0, d.hd, 64 * kc, 64, kc, 0, kc == 0);
    for (int kc = 0; kc < d.th / 64; ++kc) b.stage(pi.seg_w[0], d.th, 0, d.hd, 64 * kc, 64, kc, d.hd, kc == 0);
    b.end_pass();
    b.stage(pi.dir_w[0], ld0, 0, d.hd, d.th, d.xtra_dim, 0, 0, false);
    b.end_pass();
    fold(pi.dir_b[0], -1, 0, 0, 0, 0, d.hd, d.off_hbias);
    fold(pi.seg_b[0], -1, 0, 0, 0, 0, d.hd, d.off_hbias + d.hd);
    for (int i = 1; i < 4; ++i) {
      for (int kc = 0; kc < d.hd / 64; ++kc) b.stage(pi.dir_w[i], d.hd, 0, d.hd, 64 * kc, 64, kc, 0, kc == 0);
      for (int kc = 0; kc < d.hd / 64; ++kc) b.stage(pi.seg_w[i], d.hd, 0, d.hd, 64 * kc, 64, d.hd / 64 + kc, d.hd, kc == 0);
      b.end_pass();
      fold(pi.dir_b[i], -1, 0, 0, 0, 0, d.hd, d.off_hbias + i * 2 * d.hd);
      fold(pi.seg_b[i], -1, 0, 0, 0, 0, d.hd, d.off_hbias + i * 2 * d.hd + d.hd);
    }
    // output layer: rows 0-2 = fc_rgb over the dir hidden (chunks 0..), rows 3-14 = fc_seg over the seg hidden
    for (int kc = 0; kc < 2 * d.hd / 64; ++kc) {
      const bool is_rgb = kc < d.hd / 64;
      b.stage(is_rgb ? pi.rgb_w : pi.segf_w, d.hd, 0, 16, 64 * (is_rgb ? kc : kc - d.hd / 64), 64, kc, 0, kc == 0);
      PackSrc& src = hp.pack[b.ns - 1].src[0];
      src.dst_row0 = is_rgb ? 0 : 3;
      src.nrows = is_rgb ? 3 : 12;
    }
    b.end_pass();
    fold(pi.rgb_b, -1, 0, 0, 0, 0, 3, d.off_outb);
    fold(pi.segf_b, -1, 0, 0, 0, 0, 12, d.off_outb + 3);
  }
  hp.plan.num_stages = b.ns;
  hp.plan.total_bytes = (int32_t)b.off;
  finalize_plan(hp.plan);
  return SAHS_OK;
}

int sahs_build_bwd_plan(const sahs_model_spec& s, const float* const* params, HostPlan& hp) {
  NetDims& d = hp.dims;
  int rc = sahs_make_dims(s, d, true);
  if (rc) {
    sahs_set_error("unsupported model spec (dims check %d)", rc);
    return SAHS_EUNSUPPORTED;
  }
  ParamIndex pi;
  if (!index_params(s, pi)) return SAHS_EINVAL;
  Builder b{hp, params};
  b.f16 = true;    // fp16 operands; the caller scales d_raw into fp16's range (see field_bwd.cu)
  hp.num_fold = 0;
  hp.num_copy = 0;
  const int CW = SAHS_DRIVING_DIM + SAHS_POSE_CODE_DIM;
  const int hd = d.hd, th = d.th;
  // output layer: dOUT (chunk 0, cols 0-2 rgb, 3-14 seg) -> d(dir hidden) | d(seg hidden)
  b.tstage(pi.rgb_w, hd, 0, 3, 0, 0, hd, hd, 0, 0, true);
  b.tstage(pi.segf_w, hd, 0, 12, 3, 0, hd, hd, 0, hd, true);
  b.end_pass();
  for (int i = 3; i >= 1; --i) {
    for (int kc = 0; kc < hd / 64; ++kc) b.tstage(pi.dir_w[i], hd, 64 * kc, 64, 0, 0, hd, hd, kc, 0, kc == 0);
    for (int kc = 0; kc < hd / 64; ++kc) b.tstage(pi.seg_w[i], hd, 64 * kc, 64, 0, 0, hd, hd, hd / 64 + kc, hd, kc == 0);
    b.end_pass();
  }
  const int ld0 = th + d.xtra_dim;
  for (int kc = 0; kc < hd / 64; ++kc) b.tstage(pi.dir_w[0], ld0, 64 * kc, 64, 0, th, 64, d.xtra_dim, kc, 0, kc == 0);
  b.end_pass();   // d[PE(dir) | embedding]
  for (int half = 0; half < th / 128; ++half) {
    for (int kc = 0; kc < hd / 64; ++kc) b.tstage(pi.dir_w[0], ld0, 64 * kc, 64, 0, 128 * half, 128, 128, kc, 128 * half, kc == 0);
    for (int kc = 0; kc < hd / 64; ++kc) b.tstage(pi.seg_w[0], th, 64 * kc, 64, 0, 128 * half, 128, 128, hd / 64 + kc, 128 * half, false);
  }
  b.end_pass();   // d feat
  for (int half = 0; half < th / 128; ++half)
    for (int kc = 0; kc < th / 64; ++kc) b.tstage(pi.feat_w, th, 64 * kc, 64, 0, 128 * half, 128, 128, kc, 128 * half, kc == 0);
  b.end_pass();
  const int tin = d.e1_dim + d.ct_len;
  for (int i = d.t_layers - 1; i >= 0; --i) {
    const bool first = i == 0, skip = i == d.t_skip;
    const int ld = first ? tin : (skip ? th + tin : th);
    if (first || skip) {
      for (int kc = 0; kc < th / 64; ++kc)
        b.tstage(pi.trunk_w[i], ld, 64 * kc, 64, 0, first ? 0 : th, d.e1_k, d.e1_dim, kc, 0, kc == 0);
      b.end_pass();   // d encoding
    }
    if (!first) {
      for (int half = 0; half < th / 128; ++half)
        for (int kc = 0; kc < th / 64; ++kc) b.tstage(pi.trunk_w[i], ld, 64 * kc, 64, 0, 128 * half, 128, 128, kc, 128 * half, kc == 0);
      b.end_pass();
    }
  }
  if (d.use_w) {
    const int in0 = d.e0_dim + CW;
    for (int i = d.w_layers - 1; i >= 1; --i) {
      const bool skip = i == d.w_skip;
      const int ldw = skip ? d.wh + in0 : d.wh, ldh = skip ? d.hh + in0 : d.hh;
      for (int kc = 0; kc < d.wh / 64; ++kc) b.tstage(pi.warp_w[i], ldw, 64 * kc, 64, 0, 0, d.wh, d.wh, kc, 0, kc == 0);
      for (int kc = 0; kc < d.hh / 64; ++kc) b.tstage(pi.hyp_w[i], ldh, 64 * kc, 64, 0, 0, d.hh, d.hh, d.wh / 64 + kc, d.wh, kc == 0);
      b.end_pass();
    }
  }
  hp.plan.num_stages = b.ns;
  hp.plan.total_bytes = (int32_t)b.off;
  finalize_plan(hp.plan);
  return SAHS_OK;
}

// --------------------------------------------------------------------------------------------------------
// pack kernel: one block per stage, fp32 -> fp16 / bf16 (PackStage::f16), written at the 128B-swizzled offset
// --------------------------------------------------------------------------------------------------------
constexpr int kPackBatch = kMaxStages;   // one launch per image: 160 x 96 B of kernel parameters (limit 32 KB since CUDA 12.1)
constexpr int kPackSlices = 8;            // CTAs per stage image (a stage is at most 256 x 64 elements)
struct PackBatch {
  PackStage st[kPackBatch];
};

__global__ void pack_stage_kernel(const __grid_constant__ PackBatch batch, uint8_t* __restrict__ out) {
  const PackStage& ps = batch.st[blockIdx.x];
  const int kshift = ps.wide ? 5 : 6;            // image columns: 32 (ST_WIDE) or 64
  const int total = ps.n << kshift;
  for (int e = blockIdx.y * blockDim.x + threadIdx.x; e < total; e += blockDim.x * gridDim.y) {
    int row = e >> kshift, col = e & ((1 << kshift) - 1);
    float v = 0.f;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const PackSrc& s = ps.src[q];
      if (s.w && row >= s.dst_row0 && row < s.dst_row0 + s.nrows && col >= s.dst_col0 && col < s.dst_col0 + s.ncols)
        v = s.transpose ? s.w[(size_t)(s.src_row0 + col - s.dst_col0) * s.ld + s.src_col0 + (row - s.dst_row0)]
                        : s.w[(size_t)(s.src_row0 + row - s.dst_row0) * s.ld + s.src_col0 + col - s.dst_col0];
    }
    if (ps.f16) {
      __half h = __float2half_rn(v);
      if (ps.lo) h = __float2half_rn(v - __half2float(h));
      *reinterpret_cast<__half*>(out + ps.dst_off + (ps.wide ? sw64_offset(row, col) : sw128_offset(row, col))) = h;
    }
    else
      *reinterpret_cast<__nv_bfloat16*>(out + ps.dst_off + (ps.wide ? sw64_offset(row, col) : sw128_offset(row, col))) =
          __float2bfloat16_rn(v);
  }
}

// grid [1,C,D,H,W] -> channel-last [D,H,W,C]
__global__ void grid_channel_last_kernel(const float* __restrict__ g, float* __restrict__ out, int C, int vox) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= C * vox) return;
  int v = idx / C, c = idx - v * C;
  out[idx] = g[(size_t)c * vox + v];
}

struct FoldBatch {
  FoldSection f[64];
  CopySection c[8];
  int nf, nc;
};

// one warp per output element of a folded section; plain sections are copied
__global__ void fold_frame_kernel(const __grid_constant__ FoldBatch fb, const float* __restrict__ driving,
                                  const float* __restrict__ pose_code, float* __restrict__ out) {
  __shared__ float cvec[SAHS_DRIVING_DIM + SAHS_POSE_CODE_DIM];
  for (int i = threadIdx.x; i < SAHS_DRIVING_DIM + SAHS_POSE_CODE_DIM; i += blockDim.x)
    cvec[i] = i < SAHS_DRIVING_DIM ? driving[i] : pose_code[i - SAHS_DRIVING_DIM];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  if ((int)blockIdx.x < fb.nf) {
    const FoldSection& f = fb.f[blockIdx.x];
    for (int o = warp; o < f.n; o += nwarps) {
      float acc = 0.f;
      if (f.w)
        for (int j = lane; j < f.ncols; j += 32) acc += f.w[(size_t)o * f.ld + f.col0 + j] * cvec[f.c_off + j];
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
      if (lane == 0) out[f.dst + o] = f.bias[o] + acc;
    }
  } else {
    const CopySection& c = fb.c[blockIdx.x - fb.nf];
    for (int i = threadIdx.x; i < c.count; i += blockDim.x) out[c.dst + i] = c.src[i];
  }
}

extern "C" int sahs_field_sizes(const sahs_model_spec* spec, size_t* packed_bytes, size_t* frame_const_bytes,
                                size_t* grid_bytes) {
  SAHS_CHECK_ARG(spec, "null spec");
  static thread_local HostPlan hp;
  int rc = sahs_build_host_plan(*spec, nullptr, hp);
  if (rc) return rc;
  if (packed_bytes) *packed_bytes = (size_t)hp.plan.total_bytes;
  if (frame_const_bytes) *frame_const_bytes = (size_t)hp.dims.fc_total * sizeof(float);
  if (grid_bytes)
    *grid_bytes = spec->use_grid ? (size_t)SAHS_GRID_CH * SAHS_GRID_RES * SAHS_GRID_RES * SAHS_GRID_RES * 4 : 0;
  return SAHS_OK;
}

extern "C" int sahs_pack_params(const sahs_model_spec* spec, int level, const float* const* params,
                                void* packed_out, float* grid_out, void* stream) {
  SAHS_CHECK_ARG(spec && params && packed_out, "null pointer");
  SAHS_CHECK_ARG(level == 0 || level == 1, "level must be 0 (coarse) or 1 (fine)");
  static thread_local HostPlan hp;
  int rc = sahs_build_host_plan(*spec, params, hp);
  if (rc) return rc;
  for (int i = 0; i < hp.plan.num_stages; ++i)
    SAHS_CHECK_ARG(hp.pack[i].src[0].w != nullptr, "a required parameter pointer is NULL");
  cudaStream_t st = (cudaStream_t)stream;
  for (int s0 = 0; s0 < hp.plan.num_stages; s0 += kPackBatch) {
    PackBatch pb;
    int cnt = hp.plan.num_stages - s0 < kPackBatch ? hp.plan.num_stages - s0 : kPackBatch;
    memcpy(pb.st, hp.pack + s0, sizeof(PackStage) * cnt);
    pack_stage_kernel<<<dim3(cnt, kPackSlices), 256, 0, st>>>(pb, (uint8_t*)packed_out);
    SAHS_LAUNCH_CHECK();
  }
  if (spec->use_grid && grid_out) {
    SAHS_CHECK_ARG(params[0] != nullptr, "spatial_embeddings pointer is NULL");
    const int vox = SAHS_GRID_RES * SAHS_GRID_RES * SAHS_GRID_RES;
    grid_channel_last_kernel<<<(SAHS_GRID_CH * vox + 255) / 256, 256, 0, st>>>(params[0], grid_out, SAHS_GRID_CH, vox);
    SAHS_LAUNCH_CHECK();
  }
  return SAHS_OK;
}

extern "C" int sahs_fold_frame(const sahs_model_spec* spec, int level, const float* const* params,
                               const float* driving, const float* pose_code, float* frame_const_out, void* stream) {
  SAHS_CHECK_ARG(spec && params && driving && pose_code && frame_const_out, "null pointer");
  SAHS_CHECK_ARG(level == 0 || level == 1, "level must be 0 (coarse) or 1 (fine)");
  static thread_local HostPlan hp;
  int rc = sahs_build_host_plan(*spec, params, hp);
  if (rc) return rc;
  FoldBatch fb;
  memcpy(fb.f, hp.fold, sizeof(FoldSection) * hp.num_fold);
  memcpy(fb.c, hp.copy, sizeof(CopySection) * hp.num_copy);
  fb.nf = hp.num_fold;
  fb.nc = hp.num_copy;
  for (int i = 0; i < fb.nf; ++i) SAHS_CHECK_ARG(fb.f[i].bias != nullptr, "a required bias pointer is NULL");
  fold_frame_kernel<<<fb.nf + fb.nc, 256, 0, (cudaStream_t)stream>>>(fb, driving, pose_code, frame_const_out);
  SAHS_LAUNCH_CHECK();
  return SAHS_OK;
}

// Host-only introspection of the plan (no device work): used by the CPU test-suite to check the stage table and
// the pack/fold descriptors against the reference layer shapes.  `params` entries are treated as opaque ids.
// Per stage 14 ints: n, ksteps, flags, a_chunk, d_col, dst_off, param_id, src_row0, src_col0, dst_row0, nrows, ncols,
// a_chunk2 (255 = none), lo.
// Per fold section 8 ints: bias_id, w_id (0 = none), ld, col0, ncols, c_off, n, dst.  Per copy 3 ints: id, count, dst.
extern "C" int sahs_debug_plan(const sahs_model_spec* spec, int param_count, int32_t* stages, int max_stages,
                               int32_t* folds, int max_folds, int32_t* copies, int max_copies, int32_t* dims_out) {
  SAHS_CHECK_ARG(spec && stages && folds && copies && dims_out, "null pointer");
  std::vector<const float*> ids(param_count);
  for (int i = 0; i < param_count; ++i) ids[i] = reinterpret_cast<const float*>((uintptr_t)(i + 1) * 4);
  static thread_local HostPlan hp;
  int rc = sahs_build_host_plan(*spec, ids.data(), hp);
  if (rc) return rc;
  SAHS_CHECK_ARG(hp.plan.num_stages <= max_stages && hp.num_fold <= max_folds && hp.num_copy <= max_copies,
                 "output arrays too small");
  auto id_of = [](const float* p) -> int32_t { return p ? (int32_t)((uintptr_t)p / 4) : 0; };
  for (int i = 0; i < hp.plan.num_stages; ++i) {
    const StageRec& r = hp.plan.st[i];
    const PackStage& ps = hp.pack[i];
    int32_t* o = stages + 14 * i;
    o[0] = r.n8 * 8; o[1] = r.kflags & 7; o[2] = r.kflags >> 3; o[3] = r.a_chunk | (r.a_k16 << 8); o[4] = r.d_col8 * 8;
    o[5] = (int32_t)ps.dst_off; o[6] = id_of(ps.src[0].w); o[7] = ps.src[0].src_row0; o[8] = ps.src[0].src_col0;
    o[9] = ps.src[0].dst_row0; o[10] = ps.src[0].nrows; o[11] = ps.src[0].ncols; o[12] = r.a_chunk2; o[13] = ps.lo;
  }
  for (int i = 0; i < hp.num_fold; ++i) {
    const FoldSection& f = hp.fold[i];
    int32_t* o = folds + 8 * i;
    o[0] = id_of(f.bias); o[1] = id_of(f.w); o[2] = f.ld; o[3] = f.col0; o[4] = f.ncols; o[5] = f.c_off; o[6] = f.n;
    o[7] = f.dst;
  }
  for (int i = 0; i < hp.num_copy; ++i) {
    int32_t* o = copies + 3 * i;
    o[0] = id_of(hp.copy[i].src); o[1] = hp.copy[i].count; o[2] = hp.copy[i].dst;
  }
  const NetDims& d = hp.dims;
  int32_t dd[] = {hp.plan.num_stages, hp.num_fold, hp.num_copy, hp.plan.total_bytes, d.fc_total, d.e0_dim, d.e0_k,
                  d.e1_dim, d.e1_k, d.e0_resident, d.e0_chunk_base, d.whh, d.off_wbias, d.off_wfinal, d.off_tbias,
                  d.off_featb, d.off_alpha, d.off_hbias, d.off_outb, d.xtra_dim, d.w_split};
  memcpy(dims_out, dd, sizeof(dd));
  return SAHS_OK;
}


// ---- training support ----------------------------------------------------------------------------------------------
extern "C" int sahs_train_layout(const sahs_model_spec* spec, int32_t* out, int max_out) {
  SAHS_CHECK_ARG(spec && out && max_out >= 40, "need room for 40 ints");
  NetDims d;
  int rc = sahs_make_dims(*spec, d, true);
  if (rc) { sahs_set_error("unsupported model spec (dims check %d)", rc); return SAHS_EUNSUPPORTED; }
  static thread_local HostPlan hp;
  rc = sahs_build_host_plan(*spec, nullptr, hp, true);
  if (rc) return rc;
  const int32_t train_packed_bytes = hp.plan.total_bytes;
  rc = sahs_build_bwd_plan(*spec, nullptr, hp);
  if (rc) return rc;
  int32_t v[40] = {d.tx_e0, d.e0_k, d.tx_wh, d.whh, d.w_layers, d.tx_e1, d.e1_k, d.tx_th, d.th, d.t_layers, d.tx_feat,
                   d.tx_xtra, d.tx_hh, d.tx_total, d.td_wh, d.td_final, d.td_th, d.td_feat, d.td_hh, d.td_out, d.td_total,
                   d.n_mask_layers, d.e0_dim, d.e1_dim, d.xtra_dim, d.wh, d.hh, d.w_skip, d.t_skip, d.ct_off, d.ct_len,
                   d.use_w, d.hd, hp.plan.total_bytes, hp.plan.num_stages, d.fc_total, train_packed_bytes, 0, 0, 0};
  memcpy(out, v, sizeof(v));
  return SAHS_OK;
}

extern "C" int sahs_pack_params_train(const sahs_model_spec* spec, int level, const float* const* params,
                                      void* packed_out, void* stream) {
  SAHS_CHECK_ARG(spec && params && packed_out, "null pointer");
  SAHS_CHECK_ARG(level == 0 || level == 1, "level must be 0 (coarse) or 1 (fine)");
  static thread_local HostPlan hp;
  int rc = sahs_build_host_plan(*spec, params, hp, true);
  if (rc) return rc;
  for (int s0 = 0; s0 < hp.plan.num_stages; s0 += kPackBatch) {
    PackBatch pb;
    int cnt = hp.plan.num_stages - s0 < kPackBatch ? hp.plan.num_stages - s0 : kPackBatch;
    memcpy(pb.st, hp.pack + s0, sizeof(PackStage) * cnt);
    pack_stage_kernel<<<dim3(cnt, kPackSlices), 256, 0, (cudaStream_t)stream>>>(pb, (uint8_t*)packed_out);
    SAHS_LAUNCH_CHECK();
  }
  return SAHS_OK;
}

extern "C" int sahs_pack_params_bwd(const sahs_model_spec* spec, int level, const float* const* params,
                                    void* packed_t_out, void* stream) {
  SAHS_CHECK_ARG(spec && params && packed_t_out, "null pointer");
  SAHS_CHECK_ARG(level == 0 || level == 1, "level must be 0 (coarse) or 1 (fine)");
  static thread_local HostPlan hp;
  int rc = sahs_build_bwd_plan(*spec, params, hp);
  if (rc) return rc;
  for (int i = 0; i < hp.plan.num_stages; ++i)
    SAHS_CHECK_ARG(hp.pack[i].src[0].w != nullptr, "a required parameter pointer is NULL");
  for (int s0 = 0; s0 < hp.plan.num_stages; s0 += kPackBatch) {
    PackBatch pb;
    int cnt = hp.plan.num_stages - s0 < kPackBatch ? hp.plan.num_stages - s0 : kPackBatch;
    memcpy(pb.st, hp.pack + s0, sizeof(PackStage) * cnt);
    pack_stage_kernel<<<dim3(cnt, kPackSlices), 256, 0, (cudaStream_t)stream>>>(pb, (uint8_t*)packed_t_out);
    SAHS_LAUNCH_CHECK();
  }
  return SAHS_OK;
}
