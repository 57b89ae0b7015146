// (4) hierarchical importance resampling + merge: one warp per ray.
// ref: nerf/nerf_helpers.py:454-497 (sample_pdf_2), call site nerf/train_utils.py:157-166.
//
// Bit-exactness contract: fed the reference's weights, the indices returned by
// searchsorted must equal the reference's on CPU.  That requires reproducing ATen's CPU reduction orders:
//   * sum(w): vectorized_inner_sum -- 8 fp32 lanes, 4 interleaved accumulators over whole 8-wide vectors,
//     remaining vectors into accumulator 0, accumulators folded 1..3 into 0, scalar tail summed first into a
//     scalar, then the 8 lanes added to it in lane order.  No FMA.
//   * pdf = w / sum: IEEE division.
//   * cumsum: double accumulator, rounded to fp32 per element.  For this domain (pdf in [~1e-5/(1+eps), 1],
//     <= 254 terms) every partial double sum is exact (<= 24+17+8 significant bits), so a parallel scan in
//     double gives the same bits as ATen's sequential loop.
//   * z_mid, t, sample: separate fp32 mul/add (no contraction).
#include "sahs_common.cuh"

namespace {

constexpr int kWarps = 8;
constexpr int kMaxS = 256;      // coarse samples per ray
constexpr int kMaxMerged = 512; // S + n_fine

constexpr int kBlocksPerSm = 4; // 32 resident warps per SM (6 blocks = 40 registers measured no faster: issue bound)

// per-warp scratch in dynamic shared memory, sized for the launch's S and NF (so short rays leave room for more warps)
struct WarpScratch {
  float* w;        // w[j] = weights[j+1] + 1e-5, j < n = S-2          [S]
  float* cdf;      // n+1 entries                                      [S]
  float* bins;     // S-1 mid points; reused for the 64 sorted samples [max(S, 64)]
  float* merged;   // cat(z, samples), padded to a power of two        [P]
};
__host__ __device__ inline int merged_pow2(int S, int NF) {
  int P = 1;
  while (P < S + NF) P <<= 1;
  return P;
}
__host__ __device__ inline int scratch_floats(int S, int NF) { return 2 * S + (S > 64 ? S : 64) + merged_pow2(S, NF); }

__device__ __forceinline__ double shfl_up_double(double v, int o) {
  int lo = __double2loint(v), hi = __double2hiint(v);
  lo = __shfl_up_sync(0xffffffffu, lo, o);
  hi = __shfl_up_sync(0xffffffffu, hi, o);
  return __hiloint2double(hi, lo);
}

__global__ void __launch_bounds__(kWarps * 32, kBlocksPerSm)
sample_pdf_merge_kernel(const float* __restrict__ z, const float* __restrict__ bins_in,
                        const float* __restrict__ weights, const float* __restrict__ u, int u_per_ray, RngArg rng,
                        int R, int S, int NF, float* __restrict__ z_samples, float* __restrict__ z_merged,
                        int64_t* __restrict__ inds_out) {
  extern __shared__ float scratch_all[];
  WarpScratch sm;
  sm.w = scratch_all + (size_t)(threadIdx.x >> 5) * scratch_floats(S, NF);
  sm.cdf = sm.w + S;
  sm.bins = sm.cdf + S;
  sm.merged = sm.bins + (S > 64 ? S : 64);
  const int lane = threadIdx.x & 31;
  const int n = S - 2;        // number of pdf entries
  const int nb = S - 1;       // number of bins == number of cdf entries
  const int warps_total = gridDim.x * kWarps;
  for (int r = blockIdx.x * kWarps + (threadIdx.x >> 5); r < R; r += warps_total) {
    if (bins_in) {
      // public sample_pdf_2(bins, weights, ...) form: bins [R,S-1], weights [R,S-2], no merge
      for (int j = lane; j < n; j += 32) sm.w[j] = __fadd_rn(weights[(size_t)r * n + j], 1e-5f);
      for (int j = lane; j < nb; j += 32) sm.bins[j] = bins_in[(size_t)r * nb + j];
    } else {
      const float* zr = z + (size_t)r * S;
      const float* wr = weights + (size_t)r * S;
      for (int j = lane; j < n; j += 32) sm.w[j] = __fadd_rn(wr[j + 1], 1e-5f);
      for (int j = lane; j < nb; j += 32) sm.bins[j] = __fmul_rn(0.5f, __fadd_rn(zr[j + 1], zr[j]));
      for (int j = lane; j < S; j += 32) sm.merged[j] = zr[j];
    }
    __syncwarp();
    // ---- total in ATen's vectorized_inner_sum order -------------------------------------------
    const int nv = n / 8, size_ilp = nv / 4;
    float p0 = 0.f;
    if (lane < 8) {
      float p1 = 0.f, p2 = 0.f, p3 = 0.f;
      for (int i = 0; i < size_ilp; ++i) {
        p0 = __fadd_rn(p0, sm.w[(4 * i + 0) * 8 + lane]);
        p1 = __fadd_rn(p1, sm.w[(4 * i + 1) * 8 + lane]);
        p2 = __fadd_rn(p2, sm.w[(4 * i + 2) * 8 + lane]);
        p3 = __fadd_rn(p3, sm.w[(4 * i + 3) * 8 + lane]);
      }
      for (int i = size_ilp * 4; i < nv; ++i) p0 = __fadd_rn(p0, sm.w[i * 8 + lane]);
      p0 = __fadd_rn(p0, p1);
      p0 = __fadd_rn(p0, p2);
      p0 = __fadd_rn(p0, p3);
    }
    float total = 0.f;
    for (int k = nv * 8; k < n; ++k) total = __fadd_rn(total, sm.w[k]);
#pragma unroll
    for (int l = 0; l < 8; ++l) total = __fadd_rn(total, __shfl_sync(0xffffffffu, p0, l));
    // ---- pdf and cdf ---------------------------------------------------------------------------
    // lane owns the contiguous slice [lane*per, lane*per+per) of the pdf
    const int per = (n + 31) / 32;
    double local = 0.0;
    for (int q = 0; q < per; ++q) {
      int j = lane * per + q;
      if (j < n) local += (double)__fdiv_rn(sm.w[j], total);
    }
    double incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      double up = shfl_up_double(incl, o);
      if (lane >= o) incl += up;
    }
    double run = incl - local;  // exclusive prefix of this lane's slice
    if (lane == 0) sm.cdf[0] = 0.f;
    for (int q = 0; q < per; ++q) {
      int j = lane * per + q;
      if (j < n) {
        run += (double)__fdiv_rn(sm.w[j], total);
        sm.cdf[j + 1] = (float)run;
      }
    }
    __syncwarp();
    // ---- invert the cdf --------------------------------------------------------------------------
    for (int i = lane; i < NF; i += 32) {
      const float uu = rng.on ? rng_uniform(rng, (uint64_t)r * NF + i) : (u_per_ray ? u[(size_t)r * NF + i] : u[i]);
      // searchsorted(right=True): first index with cdf[idx] > u  == number of entries <= u
      int lo = 0, hi = nb;
      while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (sm.cdf[mid] <= uu) lo = mid + 1; else hi = mid;
      }
      const int ind = lo;
      const int below = max(ind - 1, 0), above = min(ind, nb - 1);
      const float cb = sm.cdf[below], ca = sm.cdf[above];
      const float bb = sm.bins[below], ba = sm.bins[above];
      float denom = __fsub_rn(ca, cb);
      if (denom < 1e-5f) denom = 1.0f;
      const float t = __fdiv_rn(__fsub_rn(uu, cb), denom);
      const float smp = __fadd_rn(bb, __fmul_rn(t, __fsub_rn(ba, bb)));
      z_samples[(size_t)r * NF + i] = smp;
      if (!bins_in) sm.merged[S + i] = smp;
      if (inds_out) inds_out[(size_t)r * NF + i] = ind;
    }
    __syncwarp();
    if (bins_in) continue;
    // ---- sort(cat(z, z_samples)) (values only, so any correct sort reproduces torch.sort) ---------
    const int tot = S + NF;
    // Fast path (the shapes the renderer uses): the coarse depths are already ascending, so sort the 64 new samples in
    // registers (bitonic network over warp shuffles, two values per lane) and merge the two sorted lists by rank:
    // position(z_i) = i + #{samples < z_i}, position(s_j) = j + #{coarse <= s_j}.
    bool fast = (NF == 64);
    if (fast) {
      bool ok = true;
      for (int j = lane; j + 1 < S; j += 32) ok = ok && !(sm.merged[j] > sm.merged[j + 1]);
      fast = __all_sync(0xffffffffu, ok);
    }
    if (fast) {
      float v0 = sm.merged[S + lane], v1 = sm.merged[S + 32 + lane];   // element index e = lane (v0), 32 + lane (v1)
      // Deterministic sampling (the render path: u = linspace) inverts a non-decreasing cdf at ascending u, so the
      // samples usually come out ascending already; then the 21-stage bitonic network (a quarter of the kernel's
      // instructions) is skipped.  Checked, not assumed: stochastic u, or a tie broken the other way by rounding,
      // still sorts.
      const float nx0 = sm.merged[S + lane + 1], nx1 = lane < 31 ? sm.merged[S + 32 + lane + 1] : v1;
      const bool sorted = __all_sync(0xffffffffu, !(v0 > nx0) && !(v1 > nx1));
      for (int k = sorted ? 128 : 2; k <= 64; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
          if (j == 32) {            // partner is the other register of the same lane (only when k == 64: ascending)
            const float lo = fminf(v0, v1), hi = fmaxf(v0, v1);
            v0 = lo; v1 = hi;
          } else {
            const float p0 = __shfl_xor_sync(0xffffffffu, v0, j), p1 = __shfl_xor_sync(0xffffffffu, v1, j);
            const bool lower = (lane & j) == 0;                 // this lane holds the lower index of the pair
            const bool up0 = (lane & k) == 0, up1 = ((32 + lane) & k) == 0;
            v0 = (lower == up0) ? fminf(v0, p0) : fmaxf(v0, p0);
            v1 = (lower == up1) ? fminf(v1, p1) : fmaxf(v1, p1);
          }
        }
      }
      __syncwarp();
      sm.bins[lane] = v0;            // sorted samples (bins[] is dead by now and holds at least 64 floats)
      sm.bins[32 + lane] = v1;
      __syncwarp();
      float* outp = z_merged + (size_t)r * tot;
      // samples: rank among the coarse depths (first index with z > s, i.e. number of coarse <= s)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float sv = h == 0 ? v0 : v1;
        int lo = 0, hi = S;
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (sm.merged[mid] <= sv) lo = mid + 1; else hi = mid;
        }
        outp[lane + 32 * h + lo] = sv;
      }
      // coarse depths: rank among the samples (number of samples < z)
      for (int i = lane; i < S; i += 32) {
        const float zv = sm.merged[i];
        int lo = 0, hi = 64;
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (sm.bins[mid] < zv) lo = mid + 1; else hi = mid;
        }
        outp[i + lo] = zv;
      }
      __syncwarp();
      continue;
    }
    // general path: bitonic network in shared memory over the next power of two
    const int P = merged_pow2(S, NF);
    for (int j = tot + lane; j < P; j += 32) sm.merged[j] = __int_as_float(0x7f800000);  // +inf padding
    __syncwarp();
    for (int k = 2; k <= P; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = lane; i < P; i += 32) {
          int ixj = i ^ j;
          if (ixj > i) {
            float a = sm.merged[i], b = sm.merged[ixj];
            bool up = (i & k) == 0;
            if ((a > b) == up) {
              sm.merged[i] = b;
              sm.merged[ixj] = a;
            }
          }
        }
        __syncwarp();
      }
    }
    for (int j = lane; j < tot; j += 32) z_merged[(size_t)r * tot + j] = sm.merged[j];
    __syncwarp();
  }
}

// ---- the renderer's shape: 64 coarse depths, 64 new samples ---------------------------------------------------------
// Same arithmetic as above, element for element (the bit-exactness contract), restructured so that a ray costs ~1/3 of
// the warp instructions: everything unrolled for n = 62 pdf entries (two per lane, pdf kept in registers, divided once),
// a branch-free 6-probe search of the 63-entry cdf, and a merge that needs no search at all:
//   * sample j came out of bin ind_j, i.e. it lies between the mid points around z[ind_j]: #{z <= s_j} is ind_j or
//     ind_j + 1 -- the hint is corrected by comparing against z itself (any hint gives the exact count; the loops run
//     0-1 times), so nothing rests on rounding behaviour of the interpolation;
//   * with ascending samples (deterministic u: checked per ray) sample j lands at slot j + #{z <= s_j}; the slots are
//     tagged in shared memory, and every output position p then knows what it holds from one ballot: a tagged slot
//     takes its sample, an untagged one the coarse depth number (#untagged slots before p).  Four coalesced stores.
// Samples that are not ascending (stochastic u) are first sorted in registers (bitonic network over warp shuffles) and
// counted with a 7-probe search; rays whose coarse depths are not ascending sort cat(z, samples) with the bitonic network
// in shared memory.  The values written are the same whichever path runs.
constexpr int kFastBlocksPerSm = 6;
struct alignas(16) FastScratch {       // per warp
  float zs[128];           // z[0..63] | samples[64..127]  (== cat(z, z_samples): the fallback sorts it in place)
  float w[64];             // weights[1:-1] + 1e-5 (62 used)
  float cdf[64];           // 63 used
  float bins[64];          // 63 used
  uint32_t tag[128];       // slot -> 1 + index of the sample it holds, 0: a coarse depth
};

__global__ void __launch_bounds__(kWarps * 32, kFastBlocksPerSm)
sample_pdf_merge64_kernel(const float* __restrict__ z, const float* __restrict__ weights, const float* __restrict__ u,
                          int u_per_ray, RngArg rng, int R, float* __restrict__ z_samples,
                          float* __restrict__ z_merged, int64_t* __restrict__ inds_out) {
  constexpr int S = 64, NF = 64, n = 62, nb = 63;
  __shared__ FastScratch scratch[kWarps];
  FastScratch& sm = scratch[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  const uint32_t lt_mask = (1u << lane) - 1u;
  const int warps_total = gridDim.x * kWarps;
  float u0 = 0.f, u1 = 0.f;
  if (!rng.on && !u_per_ray) { u0 = u[lane]; u1 = u[32 + lane]; }
  for (int r = blockIdx.x * kWarps + (threadIdx.x >> 5); r < R; r += warps_total) {
    const float2 z2 = reinterpret_cast<const float2*>(z + (size_t)r * S)[lane];        // z[2 lane], z[2 lane + 1]
    const float2 w2 = reinterpret_cast<const float2*>(weights + (size_t)r * S)[lane];
    const float z_next = __shfl_down_sync(0xffffffffu, z2.x, 1);                       // z[2 lane + 2]
    reinterpret_cast<float2*>(sm.zs)[lane] = z2;
    sm.bins[2 * lane] = __fmul_rn(0.5f, __fadd_rn(z2.y, z2.x));
    if (lane < 31) sm.bins[2 * lane + 1] = __fmul_rn(0.5f, __fadd_rn(z_next, z2.y));
    if (lane > 0) sm.w[2 * lane - 1] = __fadd_rn(w2.x, 1e-5f);                         // w[j] = weights[j + 1] + 1e-5
    if (lane < 31) sm.w[2 * lane] = __fadd_rn(w2.y, 1e-5f);
#pragma unroll
    for (int k = 0; k < 4; ++k) sm.tag[32 * k + lane] = 0u;
    bool ascending = !(z2.x > z2.y) && (lane == 31 || !(z2.y > z_next));
    __syncwarp();
    // ---- total in ATen's vectorized_inner_sum order (n = 62: 7 vectors of 8, 6 tail elements) --------------------
    float p0 = 0.f;
    if (lane < 8) {
      p0 = sm.w[lane];                                       // accumulators start at 0: 0 + w is w
      p0 = __fadd_rn(p0, sm.w[32 + lane]);                   // vectors 4..6 go to accumulator 0
      p0 = __fadd_rn(p0, sm.w[40 + lane]);
      p0 = __fadd_rn(p0, sm.w[48 + lane]);
      p0 = __fadd_rn(p0, sm.w[8 + lane]);                    // then accumulators 1..3 are folded in
      p0 = __fadd_rn(p0, sm.w[16 + lane]);
      p0 = __fadd_rn(p0, sm.w[24 + lane]);
    }
    float total = sm.w[56];
#pragma unroll
    for (int k = 57; k < n; ++k) total = __fadd_rn(total, sm.w[k]);
#pragma unroll
    for (int l = 0; l < 8; ++l) total = __fadd_rn(total, __shfl_sync(0xffffffffu, p0, l));
    // ---- pdf (two entries per lane, in registers) and cdf (double scan, exact in this domain) -----------------------
    float pa = 0.f, pb = 0.f;
    if (lane < 31) {
      const float2 wv = reinterpret_cast<const float2*>(sm.w)[lane];
      pa = __fdiv_rn(wv.x, total);
      pb = __fdiv_rn(wv.y, total);
    }
    double local = 0.0;
    local += (double)pa;
    local += (double)pb;
    double incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double up = shfl_up_double(incl, o);
      if (lane >= o) incl += up;
    }
    double run = incl - local;
    if (lane == 0) sm.cdf[0] = 0.f;
    if (lane < 31) {
      run += (double)pa;
      sm.cdf[2 * lane + 1] = (float)run;
      run += (double)pb;
      sm.cdf[2 * lane + 2] = (float)run;
    }
    __syncwarp();
    // ---- invert the cdf: samples lane and 32 + lane ------------------------------------------------------------------
    float smp[2];
    int ind[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int i = lane + 32 * h;
      float uu = h == 0 ? u0 : u1;
      if (rng.on) uu = rng_uniform(rng, (uint64_t)r * NF + i);
      else if (u_per_ray) uu = u[(size_t)r * NF + i];
      // searchsorted(right=True) on a non-decreasing cdf: number of entries <= u, 63 = 32+16+8+4+2+1 probes cover it
      int pos = 0;
#pragma unroll
      for (int step = 32; step > 0; step >>= 1)
        if (sm.cdf[pos + step - 1] <= uu) pos += step;
      const int below = max(pos - 1, 0), above = min(pos, nb - 1);
      const float cb = sm.cdf[below], ca = sm.cdf[above];
      const float bb = sm.bins[below], ba = sm.bins[above];
      float denom = __fsub_rn(ca, cb);
      if (denom < 1e-5f) denom = 1.0f;
      const float t = __fdiv_rn(__fsub_rn(uu, cb), denom);
      smp[h] = __fadd_rn(bb, __fmul_rn(t, __fsub_rn(ba, bb)));
      ind[h] = pos;
      z_samples[(size_t)r * NF + i] = smp[h];
      if (inds_out) inds_out[(size_t)r * NF + i] = pos;
      sm.zs[S + i] = smp[h];
    }
    __syncwarp();
    float* outp = z_merged + (size_t)r * (S + NF);
    const bool z_ascending = __all_sync(0xffffffffu, ascending);
    const bool s_ascending = __all_sync(0xffffffffu, !(smp[0] > sm.zs[S + lane + 1]) &&
                                                         (lane == 31 || !(smp[1] > sm.zs[S + 32 + lane + 1])));
    if (z_ascending) {
      if (!s_ascending) {
        // stochastic u: sort the 64 samples in registers (bitonic network over warp shuffles, element lane in v0 and
        // 32 + lane in v1); the bin hints no longer belong to the sorted order, so the counts are searched
        float v0 = smp[0], v1 = smp[1];
        for (int k = 2; k <= 64; k <<= 1) {
          for (int j = k >> 1; j > 0; j >>= 1) {
            if (j == 32) {            // partner is the other register of the same lane (only when k == 64: ascending)
              const float lo = fminf(v0, v1), hi = fmaxf(v0, v1);
              v0 = lo; v1 = hi;
            } else {
              const float q0 = __shfl_xor_sync(0xffffffffu, v0, j), q1 = __shfl_xor_sync(0xffffffffu, v1, j);
              const bool lower = (lane & j) == 0;                 // this lane holds the lower index of the pair
              const bool up0 = (lane & k) == 0, up1 = ((32 + lane) & k) == 0;
              v0 = (lower == up0) ? fminf(v0, q0) : fmaxf(v0, q0);
              v1 = (lower == up1) ? fminf(v1, q1) : fmaxf(v1, q1);
            }
          }
        }
        smp[0] = v0; smp[1] = v1;
        __syncwarp();
        sm.zs[S + lane] = v0;
        sm.zs[S + 32 + lane] = v1;
#pragma unroll
        for (int h = 0; h < 2; ++h) {   // number of coarse depths <= s: 6 probes cover 63 entries, one more the 64th
          int pos = 0;
#pragma unroll
          for (int step = 32; step > 0; step >>= 1)
            if (sm.zs[pos + step - 1] <= smp[h]) pos += step;
          ind[h] = pos;
        }
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float s = smp[h];
        int c = ind[h];                                             // hint; corrected to #{coarse <= s} exactly
        while (c < S && sm.zs[c] <= s) ++c;
        while (c > 0 && sm.zs[c - 1] > s) --c;
        sm.tag[lane + 32 * h + c] = (uint32_t)(lane + 32 * h + 1);
      }
      __syncwarp();
      int coarse_before = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t tg = sm.tag[32 * k + lane];
        const uint32_t taken = __ballot_sync(0xffffffffu, tg != 0u);
        const int src = tg != 0u ? (int)(S + tg - 1u) : coarse_before + __popc(~taken & lt_mask);
        outp[32 * k + lane] = sm.zs[src];
        coarse_before += 32 - __popc(taken);
      }
    } else {
      for (int k = 2; k <= 128; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int i = lane + 32 * q, ixj = i ^ j;
            if (ixj > i) {
              const float a = sm.zs[i], b = sm.zs[ixj];
              if ((a > b) == ((i & k) == 0)) {
                sm.zs[i] = b;
                sm.zs[ixj] = a;
              }
            }
          }
          __syncwarp();
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) outp[32 * q + lane] = sm.zs[32 * q + lane];
    }
    __syncwarp();
  }
}

bool use_fast_kernel(int S, int NF) {
  if (S != 64 || NF != 64) return false;
  const char* e = getenv("SAHS_SAMPLE_PDF_GENERIC");      // measurement / test switch: the general kernel for every shape
  return !(e && e[0] == '1');
}

// dynamic shared memory of one block; above 48 KB the kernel attribute is raised once
size_t scratch_bytes(int S, int NF) {
  const size_t bytes = (size_t)kWarps * scratch_floats(S, NF) * sizeof(float);
  static size_t opted = 48 * 1024;
  if (bytes > opted) {
    cudaFuncSetAttribute(sample_pdf_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    opted = bytes;
  }
  return bytes;
}

// sample_pdf_2 + merge on (z, weights): u from the caller (shared row or per ray) or drawn in the kernel
int launch_merge(const float* z, const float* weights, const float* u, int u_per_ray, const sahs_rng* rng, int num_rays,
                 int num_samples, int num_fine, float* z_samples, float* z_merged, int64_t* inds, void* stream) {
  if (num_rays == 0) return SAHS_OK;
  SAHS_CHECK_ARG(z && weights && (u || rng) && z_samples && z_merged, "null pointer");
  SAHS_CHECK_ARG(num_samples >= 3 && num_samples <= kMaxS, "num_samples must be in [3,256]");
  SAHS_CHECK_ARG(num_fine >= 1 && num_samples + num_fine <= kMaxMerged, "num_samples + num_fine must be <= 512");
  int blocks = (num_rays + kWarps - 1) / kWarps;
  const bool aligned8 = (((uintptr_t)z | (uintptr_t)weights) & 7u) == 0;      // the fast kernel loads float2
  if (use_fast_kernel(num_samples, num_fine) && aligned8) {
    const int cap = sahs_num_sms() * kFastBlocksPerSm;
    if (blocks > cap) blocks = cap;
    sample_pdf_merge64_kernel<<<blocks, kWarps * 32, 0, (cudaStream_t)stream>>>(z, weights, u, u_per_ray, rng_arg(rng),
                                                                           num_rays, z_samples, z_merged, inds);
  } else {
    const int cap = sahs_num_sms() * kBlocksPerSm;
    if (blocks > cap) blocks = cap;
    const size_t smem = scratch_bytes(num_samples, num_fine);
    sample_pdf_merge_kernel<<<blocks, kWarps * 32, smem, (cudaStream_t)stream>>>(z, nullptr, weights, u, u_per_ray,
                                                                            rng_arg(rng), num_rays, num_samples,
                                                                            num_fine, z_samples, z_merged, inds);
  }
  SAHS_LAUNCH_CHECK();
  return SAHS_OK;
}

}  // namespace

extern "C" int sahs_sample_pdf_merge(const float* z, const float* weights, const float* u, int u_per_ray,
                                     int num_rays, int num_samples, int num_fine, float* z_samples, float* z_merged,
                                     int64_t* inds, void* stream) {
  if (num_rays == 0) return SAHS_OK;
  SAHS_CHECK_ARG(u, "null pointer");
  return launch_merge(z, weights, u, u_per_ray, nullptr, num_rays, num_samples, num_fine, z_samples, z_merged, inds,
                      stream);
}

extern "C" int sahs_sample_pdf_merge_rng(const float* z, const float* weights, const sahs_rng* rng, int num_rays,
                                         int num_samples, int num_fine, float* z_samples, float* z_merged,
                                         int64_t* inds, void* stream) {
  if (num_rays == 0) return SAHS_OK;
  SAHS_CHECK_ARG(rng, "null pointer");
  return launch_merge(z, weights, nullptr, 1, rng, num_rays, num_samples, num_fine, z_samples, z_merged, inds, stream);
}

extern "C" int sahs_sample_pdf(const float* bins, const float* weights, const float* u, int u_per_ray, int num_rays,
                               int num_bins, int num_fine, float* samples, int64_t* inds, void* stream) {
  if (num_rays == 0) return SAHS_OK;
  SAHS_CHECK_ARG(bins && weights && u && samples, "null pointer");
  SAHS_CHECK_ARG(num_bins >= 2 && num_bins + 1 <= kMaxS, "num_bins must be in [2,255]");
  SAHS_CHECK_ARG(num_fine >= 1 && num_fine <= kMaxMerged, "num_fine must be <= 512");
  if (num_rays == 0) return SAHS_OK;
  int blocks = (num_rays + kWarps - 1) / kWarps;
  int cap = sahs_num_sms() * kBlocksPerSm;
  if (blocks > cap) blocks = cap;
  const size_t smem = scratch_bytes(num_bins + 1, num_fine);
  sample_pdf_merge_kernel<<<blocks, kWarps * 32, smem, (cudaStream_t)stream>>>(nullptr, bins, weights, u, u_per_ray,
                                                                          rng_arg(nullptr), num_rays, num_bins + 1,
                                                                          num_fine, samples, nullptr, inds);
  SAHS_LAUNCH_CHECK();
  return SAHS_OK;
}
