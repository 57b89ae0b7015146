// (2c) Weight gradients of the fused field: dW_l[n,k] = sum_p dY_l[p,n] X_{l-1}[p,k], db_l[n] = sum_p dY_l[p,n],
// straight from the two training tapes with tcgen05.
//
//   * operands: the tapes are tile-major chunk images (tape[tile][slot][128 points x 64 features, 128B-swizzled], the
//     forward / backward kernels' own operand chunks, see sahs_make_dims), so a chunk is one contiguous 16 KB bulk copy
//     (cp.async.bulk) and lands in shared memory as an MN-major UMMA operand tile (features contiguous, K = points);
//   * one work unit = (<=128 output rows of one layer, <=4 input chunks, a range of tiles): the accumulator
//     D[128, 64*nB + 16] stays in TMEM across the whole tile range (the extra 16 columns multiply dY by a constant
//     "ones" operand, i.e. the bias gradient), then is added to the fp32 gradient buffers with red.global.add;
//   * 1 CTA per SM: warp 0 TMA, warp 1 MMA issue, warps 2-5 epilogue; two 96 KB operand stages.
// HBM-bound: every unit streams its dY and X boxes once (about 3 MB per tile over all layers).
#include <string.h>
#include <vector>
#include "field_dev.cuh"

namespace {

constexpr int kWThreads = 192;
constexpr int kWStages = 2;
constexpr int kWChunksPerStage = 6;                        // A0 A1 B0 B1 B2 B3
constexpr int kWStageBytes = kWChunksPerStage * kChunkBytes;   // 96 KB
constexpr int kWSmemOnes = kWStages * kWStageBytes;        // 192 KB
constexpr int kWSmemBar = kWSmemOnes + kChunkBytes;        // 208 KB
constexpr int kWSmemTotal = kWSmemBar + 128;
constexpr int kWTmemCols = 512;

struct WOut {          // where a slice of the accumulator goes
  float* w;            // fp32 [rows, ld] gradient of the weight (NULL: unused group)
  float* b;            // fp32 [rows] bias gradient (may be NULL)
  int32_t ld, m0, nrows;
  int32_t col0[4];     // destination column of input chunk j (-1: skip)
  int32_t ncols[4];    // valid columns of chunk j
};
struct WUnit {
  int32_t t0, t1;      // tile range
  int32_t a_col;       // first column of dY in the gradient tape (two 64-column chunks are loaded), multiple of 64
  int32_t nB;
  int32_t b_col[4];    // first column of each 64-wide input chunk in the activation tape, multiples of 64
  WOut g[2];
};


// MN-major SWIZZLE_128B operand: 64-feature panels of [128 points x 128 B], panels 16 KB apart (LBO), 8-point groups
// 1 KB apart (SBO); cute::UMMA::make_umma_desc<Major::MN> / DeepGEMM make_umma_desc conventions.
__device__ __forceinline__ uint64_t umma_smem_desc_mn_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(kChunkBytes >> 4) << 16;   // LBO
  d |= (uint64_t)(1024 >> 4) << 32;          // SBO
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ uint32_t umma_idesc_mn_f16_m128(uint32_t n) {   // fp16 x fp16 -> fp32, A and B MN-major
  return (1u << 4) | (1u << 15) | (1u << 16) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}

__global__ void __launch_bounds__(kWThreads, 1)
field_wgrad_kernel(const uint8_t* __restrict__ tape_x, const uint8_t* __restrict__ tape_d, int slots_x, int slots_d,
                   const WUnit* __restrict__ units, int nunits, int* __restrict__ status) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* ones = smem + kWSmemOnes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kWSmemBar);
  uint64_t* full = bars;                 // [kWStages]
  uint64_t* empty = bars + kWStages;     // [kWStages]
  uint64_t* acc_full = bars + 2 * kWStages;
  uint64_t* acc_empty = bars + 2 * kWStages + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * kWStages + 2);
  // warp index through a shuffle: the compiler then knows it is warp-uniform, keeps the role dispatch convergent and
  // the MMA issue loop on the uniform datapath (with the plain threadIdx.x >> 5 every tcgen05 instruction below gets a
  // divergence guard + R2UR moves with scoreboard waits, ~4x slower issue)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) { status[0] = 2; __trap(); }
    for (int i = 0; i < kWStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, 128);
    fence_mbar_init();
  }
  // constant "ones" operand: feature 0 of every point is 1 (bias gradient = dY^T 1)
  for (int i = threadIdx.x; i < kChunkBytes / 16; i += kWThreads) reinterpret_cast<uint4*>(ones)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  if (threadIdx.x < kTileRows) {
    const int r = threadIdx.x;
    *reinterpret_cast<__half*>(ones + sw128_offset(r, 0)) = __float2half(1.0f);
  }
  fence_proxy_async_smem();
  if (warp == 1) tmem_alloc(tmem_ptr, kWTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ================= TMA: dY (2 boxes) and X (nB boxes) of one tile per stage =================
    uint32_t stage = 0, phase = 0;
    for (int u = blockIdx.x; u < nunits; u += gridDim.x) {
      const WUnit& un = units[u];
      for (int t = un.t0; t < un.t1; ++t) {
        mbar_wait(&empty[stage], phase ^ 1, status, 100);
        if (lane == 0) {
          uint8_t* base = smem + stage * kWStageBytes;
          mbar_arrive_expect_tx(&full[stage], (uint32_t)(2 + un.nB) * kChunkBytes);
          const uint8_t* dt = tape_d + ((size_t)t * slots_d + (un.a_col >> 6)) * kChunkBytes;
          tma_bulk_g2s(base, dt, 2 * kChunkBytes, &full[stage]);                  // two adjacent dY chunks
          const uint8_t* xt = tape_x + (size_t)t * slots_x * kChunkBytes;
          for (int j = 0; j < un.nB; ++j)
            tma_bulk_g2s(base + (2 + j) * kChunkBytes, xt + (size_t)(un.b_col[j] >> 6) * kChunkBytes, kChunkBytes,
                         &full[stage]);
        }
        __syncwarp();
        if (++stage == kWStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ================= MMA: D[128, 64 nB (+16)] += dY^T [X | 1] over the tile range =================
    // whole warp converged, waits spin inside asm, elect.sync per instruction: the loop stays on the uniform datapath
    uint32_t stage = 0, phase = 0, acc_par = 0;
    int done = 0;
    const uint64_t o0 = umma_smem_desc_mn_sw128(smem_u32(ones));
    for (int u = blockIdx.x; u < nunits; u += gridDim.x) {
      const int t0 = units[u].t0, t1 = units[u].t1, nB = units[u].nB;
      if (done > 0) {   // the previous unit's accumulator must be drained
        mbar_wait_uniform<false>(acc_empty, acc_par, status, 300);
        acc_par ^= 1;
        tc_fence_after();
      }
      const uint32_t nmain = (uint32_t)nB * 64u;
      const uint32_t id_main = umma_idesc_mn_f16_m128(nmain), id_one = umma_idesc_mn_f16_m128(16);
      for (int t = t0; t < t1; ++t) {
        mbar_wait_uniform<false>(&full[stage], phase, status, 400);
        tc_fence_after();
        const uint32_t base = smem_u32(smem + stage * kWStageBytes);
        const uint64_t a0 = umma_smem_desc_mn_sw128(base);
        const uint64_t b0 = umma_smem_desc_mn_sw128(base + 2 * kChunkBytes);
        const uint32_t acc_first = (t > t0) ? 1u : 0u;
#pragma unroll
        for (uint32_t k = 0; k < 8; ++k) {   // 16 points per step = 2 KB down the panel
          const uint32_t acc = k > 0 ? 1u : acc_first;
          if (elect_one()) tc_mma_bf16(tmem_base, a0 + k * 128, b0 + k * 128, id_main, acc);
          if (elect_one()) tc_mma_bf16(tmem_base + nmain, a0 + k * 128, o0 + k * 128, id_one, acc);
        }
        if (elect_one()) tc_commit(&empty[stage]);
        if (t == t1 - 1) {
          if (elect_one()) tc_commit(acc_full);
        }
        if (++stage == kWStages) { stage = 0; phase ^= 1; }
      }
      ++done;
    }
  } else {
    // ================= epilogue: TMEM -> red.global.add into the fp32 gradient buffers =================
    const int q = warp & 3;                       // TMEM lane quarter this warp may read
    const int m = q * 32 + lane;                  // accumulator row = output feature within the unit
    const uint32_t tmem_row = tmem_base + ((uint32_t)(q * 32) << 16);
    uint32_t par = 0;
    for (int u = blockIdx.x; u < nunits; u += gridDim.x) {
      const WUnit& un = units[u];
      if (un.t1 <= un.t0) continue;
      mbar_wait(acc_full, par, status, 500);
      par ^= 1;
      tc_fence_after();
#pragma unroll 1
      for (int gi = 0; gi < 2; ++gi) {
        const WOut& g = un.g[gi];
        const bool mine = g.w != nullptr && m >= g.m0 && m < g.m0 + g.nrows;
#pragma unroll 1
        for (int j = 0; j < un.nB; ++j) {
          if (g.w == nullptr || g.col0[j] < 0) continue;     // warp-uniform
          float* dst = g.w + (size_t)(m - g.m0) * g.ld + g.col0[j];
#pragma unroll 1
          for (int c0 = 0; c0 < g.ncols[j]; c0 += 16) {
            uint32_t v[16];
            tmem_ld16(tmem_row + j * 64 + c0, v);
            tmem_ld_wait();
            if (mine) {
#pragma unroll
              for (int e = 0; e < 16; ++e)
                if (c0 + e < g.ncols[j]) atomicAdd(dst + c0 + e, __uint_as_float(v[e]));
            }
          }
        }
        if (g.w != nullptr && g.b != nullptr) {
          uint32_t v[16];
          tmem_ld16(tmem_row + un.nB * 64, v);
          tmem_ld_wait();
          if (mine) atomicAdd(g.b + (m - g.m0), __uint_as_float(v[0]));
        }
      }
      tc_fence_before();
      mbar_arrive(acc_empty);
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kWTmemCols);
  }
}

// ---- host: unit list ----------------------------------------------------------------------------------------------
struct UnitBuilder {
  std::vector<WUnit> blocks;   // one entry per (layer block); tile ranges are split later
  // chunks: list of (activation-tape column, destination column in W, valid columns)
  void add(int a_col, float* w, float* b, int ld, int m0, int nrows, const int* bcols, const int* dst, const int* valid,
           int nB, float* w2 = nullptr, float* b2 = nullptr, int ld2 = 0, int m02 = 0, int nrows2 = 0,
           const int* dst2 = nullptr, const int* valid2 = nullptr) {
    WUnit u;
    memset(&u, 0, sizeof(u));
    u.a_col = a_col;
    u.nB = nB;
    for (int j = 0; j < 4; ++j) {
      u.b_col[j] = j < nB ? bcols[j] : 0;
      u.g[0].col0[j] = j < nB ? dst[j] : -1;
      u.g[0].ncols[j] = j < nB ? valid[j] : 0;
      u.g[1].col0[j] = (j < nB && dst2) ? dst2[j] : -1;
      u.g[1].ncols[j] = (j < nB && valid2) ? valid2[j] : 0;
    }
    u.g[0].w = w; u.g[0].b = b; u.g[0].ld = ld; u.g[0].m0 = m0; u.g[0].nrows = nrows;
    u.g[1].w = w2; u.g[1].b = b2; u.g[1].ld = ld2; u.g[1].m0 = m02; u.g[1].nrows = nrows2;
    blocks.push_back(u);
  }
};

}  // namespace


// grads: device pointers to zero-initialised fp32 gradient buffers in the canonical parameter order
// (sahs_param_count entries; [0] = embedding grid, unused here).  Frame-constant input columns are left untouched
// (rank-1 terms db x cvec, done by the caller).
extern "C" int sahs_field_wgrad(const sahs_model_spec* spec, int level, float* const* grads, const void* tape_x,
                                const void* tape_d, int num_points, void* units_workspace, size_t workspace_bytes,
                                unsigned long long* upload_token, void* stream) {
  SAHS_CHECK_ARG(spec && grads && units_workspace, "null pointer");
  if (num_points == 0) return SAHS_OK;
  SAHS_CHECK_ARG(tape_x && tape_d, "null tape");
  NetDims d;
  int rc = sahs_make_dims(*spec, d, true);
  if (rc) { sahs_set_error("unsupported model spec (dims check %d)", rc); return SAHS_EUNSUPPORTED; }
  const sahs_model_spec& s = *spec;
  // canonical parameter indices (same order as field_host.cu index_params)
  int k = 1;
  int warp_w[16], warp_b[16], hyp_w[16], hyp_b[16], trunk_w[16], trunk_b[16], dir_w[4], dir_b[4], seg_w[4], seg_b[4];
  int warp_fw = -1, warp_fb = -1, hyp_fw = -1, hyp_fb = -1;
  if (s.use_warp) { for (int i = 0; i < s.warp_layers; ++i) { warp_w[i] = k++; warp_b[i] = k++; } warp_fw = k++; warp_fb = k++; }
  if (s.use_ambient) { for (int i = 0; i < s.hyper_layers; ++i) { hyp_w[i] = k++; hyp_b[i] = k++; } hyp_fw = k++; hyp_fb = k++; }
  for (int i = 0; i < s.trunk_layers; ++i) { trunk_w[i] = k++; trunk_b[i] = k++; }
  const int feat_w = k++, feat_b = k++, alpha_w = k++, alpha_b = k++;
  for (int i = 0; i < 4; ++i) { dir_w[i] = k++; dir_b[i] = k++; }
  const int rgb_w = k++, rgb_b = k++;
  for (int i = 0; i < 4; ++i) { seg_w[i] = k++; seg_b[i] = k++; }
  const int segf_w = k++, segf_b = k++;
  auto G = [&](int i) -> float* { return grads[i]; };

  UnitBuilder ub;
  const int CW = SAHS_DRIVING_DIM + SAHS_POSE_CODE_DIM;
  const int th = d.th, hd = d.hd;
  auto enc_cols = [](int base, int dim, int* bc, int* dst, int dst0, int* valid) -> int {   // 64-wide chunks of an encoding
    int n = 0;
    for (int c = 0; c < dim; c += 64) { bc[n] = base + c; dst[n] = dst0 + c; valid[n] = dim - c > 64 ? 64 : dim - c; ++n; }
    return n;
  };
  if (d.use_w) {
    const int in0 = d.e0_dim + CW;
    for (int net = 0; net < 2; ++net) {
      const int n = net == 0 ? d.wh : d.hh, lo = net == 0 ? 0 : d.wh;
      const int* wi = net == 0 ? warp_w : hyp_w;
      const int* bi = net == 0 ? warp_b : hyp_b;
      for (int i = 0; i < d.w_layers; ++i) {
        const bool first = i == 0, skip = i == d.w_skip;
        const int ld = first ? in0 : (skip ? n + in0 : n);
        const int a_col = d.td_wh + i * d.whh + lo;
        bool bias_done = false;
        if (!first) {
          int bc[4], dst[4], valid[4];
          const int nB = enc_cols(d.tx_wh + (i - 1) * d.whh + lo, n, bc, dst, 0, valid);
          ub.add(a_col, G(wi[i]), G(bi[i]), ld, 0, n, bc, dst, valid, nB);
          bias_done = true;
        }
        if (first || skip) {
          int bc[4], dst[4], valid[4];
          const int nB = enc_cols(d.tx_e0, d.e0_dim, bc, dst, first ? 0 : n, valid);
          ub.add(a_col, G(wi[i]), bias_done ? nullptr : G(bi[i]), ld, 0, n, bc, dst, valid, nB);
        }
      }
    }
    {   // fc_final (rows 0-2 of dFinal x warp h5) and fc_ambient (rows 3.. x hyper h5)
      const int h5 = d.tx_wh + (d.w_layers - 1) * d.whh;
      int bc[4] = {h5, h5 + 64, h5 + 128, 0};
      int dst_f[4] = {0, 64, -1, -1}, val_f[4] = {64, 64, 0, 0};
      int dst_a[4] = {-1, -1, 0, -1}, val_a[4] = {0, 0, 64, 0};
      ub.add(d.td_final, G(warp_fw), G(warp_fb), d.wh, 0, 3, bc, dst_f, val_f, 3, G(hyp_fw), G(hyp_fb), d.hh, 3,
             s.amb_dim, dst_a, val_a);
    }
  }
  const int tin = d.e1_dim + d.ct_len;
  for (int i = 0; i < d.t_layers; ++i) {
    const bool first = i == 0, skip = i == d.t_skip;
    const int ld = first ? tin : (skip ? th + tin : th);
    for (int half = 0; half < th / 128; ++half) {
      const int a_col = d.td_th + i * th + 128 * half;
      float* w = G(trunk_w[i]) + (size_t)128 * half * ld;
      float* b = G(trunk_b[i]) + 128 * half;
      bool bias_done = false;
      if (!first) {
        int bc[4], dst[4], valid[4];
        const int nB = enc_cols(d.tx_th + (i - 1) * th, th, bc, dst, 0, valid);
        ub.add(a_col, w, b, ld, 0, 128, bc, dst, valid, nB);
        bias_done = true;
      }
      if (first || skip) {
        int bc[4], dst[4], valid[4];
        const int nB = enc_cols(d.tx_e1, d.e1_dim, bc, dst, first ? 0 : th, valid);
        ub.add(a_col, w, bias_done ? nullptr : b, ld, 0, 128, bc, dst, valid, nB);
      }
    }
  }
  {
    int bc[4], dst[4], valid[4];
    for (int half = 0; half < th / 128; ++half) {
      const int nB = enc_cols(d.tx_th + (d.t_layers - 1) * th, th, bc, dst, 0, valid);
      ub.add(d.td_feat + 128 * half, G(feat_w) + (size_t)128 * half * th, G(feat_b) + 128 * half, th, 0, 128, bc, dst, valid, nB);
    }
    // fc_alpha: row 15 of dOUT (d sigma) x feat
    const int nB = enc_cols(d.tx_feat, th, bc, dst, 0, valid);
    ub.add(d.td_out, G(alpha_w), G(alpha_b), th, 15, 1, bc, dst, valid, nB);
    // layers_dir.0 = [feat | PE(dir) | emb], layers_seg.0 = feat
    ub.add(d.td_hh, G(dir_w[0]), G(dir_b[0]), th + d.xtra_dim, 0, hd, bc, dst, valid, nB);
    {
      int bc2[4] = {d.tx_xtra, 0, 0, 0}, dst2[4] = {th, -1, -1, -1}, val2[4] = {d.xtra_dim, 0, 0, 0};
      ub.add(d.td_hh, G(dir_w[0]), nullptr, th + d.xtra_dim, 0, hd, bc2, dst2, val2, 1);
    }
    ub.add(d.td_hh + hd, G(seg_w[0]), G(seg_b[0]), th, 0, hd, bc, dst, valid, nB);
    for (int i = 1; i < 4; ++i) {
      int n2 = enc_cols(d.tx_hh + (i - 1) * 2 * hd, hd, bc, dst, 0, valid);
      ub.add(d.td_hh + i * 2 * hd, G(dir_w[i]), G(dir_b[i]), hd, 0, hd, bc, dst, valid, n2);
      n2 = enc_cols(d.tx_hh + (i - 1) * 2 * hd + hd, hd, bc, dst, 0, valid);
      ub.add(d.td_hh + i * 2 * hd + hd, G(seg_w[i]), G(seg_b[i]), hd, 0, hd, bc, dst, valid, n2);
    }
    // output layer: rows 0-2 (rgb) x dir hidden 3, rows 3-14 (seg) x seg hidden 3
    const int h3 = d.tx_hh + 3 * 2 * hd;
    int bco[4] = {h3, h3 + 64, h3 + hd, h3 + hd + 64};
    int dst_r[4] = {0, 64, -1, -1}, val_r[4] = {64, 64, 0, 0};
    int dst_s[4] = {-1, -1, 0, 64}, val_s[4] = {0, 0, 64, 64};
    ub.add(d.td_out, G(rgb_w), G(rgb_b), hd, 0, 3, bco, dst_r, val_r, 4, G(segf_w), G(segf_b), hd, 3, 12, dst_s, val_s);
  }
  // split every block's tile range so that all SMs get work of similar size
  const int ntiles = (num_points + kTileRows - 1) / kTileRows;
  const int nsm = sahs_num_sms();
  long long total = 0;
  for (auto& b : ub.blocks) total += (long long)(2 + b.nB) * ntiles;
  const long long target = total / (2LL * nsm) + 1;          // ~2 units per SM
  std::vector<WUnit> units;
  for (auto& b : ub.blocks) {
    const long long work = (long long)(2 + b.nB) * ntiles;
    int pieces = (int)((work + target - 1) / target);
    if (pieces < 1) pieces = 1;
    if (pieces > ntiles) pieces = ntiles;
    for (int p = 0; p < pieces; ++p) {
      WUnit u = b;
      u.t0 = (int)((long long)ntiles * p / pieces);
      u.t1 = (int)((long long)ntiles * (p + 1) / pieces);
      if (u.t1 > u.t0) units.push_back(u);
    }
  }
  SAHS_CHECK_ARG(units.size() * sizeof(WUnit) <= workspace_bytes, "units workspace too small (need 256 KB)");
  cudaStream_t st = (cudaStream_t)stream;
  {
    // The unit list depends only on the spec, the point count and the gradient pointers.  Whether the workspace still
    // holds it is the CALLER's knowledge (it owns the memory): `upload_token` carries the fingerprint of the list last
    // uploaded into this workspace, and the caller zeroes it when the workspace is reallocated.  (A process-global
    // cache keyed on the workspace address would go stale when the allocator hands the address to another tensor.)
    unsigned long long h = 1469598103934665603ull;   // FNV-1a over the list
    const unsigned char* bytes = reinterpret_cast<const unsigned char*>(units.data());
    for (size_t i = 0; i < units.size() * sizeof(WUnit); ++i) h = (h ^ bytes[i]) * 1099511628211ull;
    h ^= (unsigned long long)units.size() << 1;
    if (h == 0) h = 1;
    if (!upload_token || *upload_token != h) {
      SAHS_CUDA(cudaMemcpyAsync(units_workspace, units.data(), units.size() * sizeof(WUnit), cudaMemcpyHostToDevice, st));
      SAHS_CUDA(cudaStreamSynchronize(st));   // pageable source: the vector dies with this call
      if (upload_token) *upload_token = h;
    }
  }
  for (auto& u : units) {
    bool ok = (u.a_col & 63) == 0;
    for (int j = 0; j < u.nB; ++j) ok = ok && (u.b_col[j] & 63) == 0;
    SAHS_CHECK_ARG(ok, "tape items must start at multiples of 64 columns");
  }
  SAHS_CUDA(cudaFuncSetAttribute(field_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWSmemTotal));
  int* status = sahs_status_words(2);
  SAHS_CHECK_ARG(status, "cannot allocate the diagnostic word");
  int grid = nsm < (int)units.size() ? nsm : (int)units.size();
  field_wgrad_kernel<<<grid, kWThreads, kWSmemTotal, st>>>((const uint8_t*)tape_x, (const uint8_t*)tape_d,
                                                          d.tx_total / 64, d.td_total / 64,
                                                          (const WUnit*)units_workspace, (int)units.size(), status);
  SAHS_LAUNCH_CHECK();
  return SAHS_OK;
}
