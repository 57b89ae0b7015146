// Stage-II SPADE generator (ref: nerf/_init_spade.py:114-139, :183-199, :235-373): every 3x3
// convolution of the network as ONE implicit-GEMM tcgen05 kernel, plus the two small HBM-bound helpers around it
// (instance-norm statistics, 2x2 average pooling).  Activations live in HBM as NHWC fp16.
//
//   D[128 output pixels, N <= 128 output channels] = sum over (tap, 64-channel chunk) A_tap[128 x 64] * W_tap[N x 64]^T
//
//   * A operand: gathered by the worker threads -- one thread per output pixel reads the 128 contiguous bytes of its
//     input pixel for the current tap (zero outside the image) and writes them as one row of a K-major SWIZZLE_128B
//     operand tile.  The gather coordinate function is the only thing that distinguishes the layer types of the network:
//     stride-1 conv, stride-2 conv (residual_downsample), stride-2 transposed conv (residual_upsample: input pixel
//     (o + 1 - k) / 2 when even), and a nearest-neighbour resize of the input by a power of two in either direction
//     (F.interpolate(fid, size=x.size()), nn.Upsample(scale_factor=2)): resized / upsampled tensors are never stored.
//     The two 3-channel convs take all nine taps in one K chunk (9 x 4 channels).
//   * B operand: the layer's weights, pre-packed per (N tile, tap, chunk) in the same swizzled layout and streamed by the
//     TMA unit (cp.async.bulk) through a 3-slot ring; accumulators in TMEM (128 columns per CTA, two CTAs per SM).
//   * epilogue (same worker threads, TMEM -> registers -> NHWC): bias, optional ReLU, optional residual add; or the whole
//     SPADE modulation: the N tile holds [gamma(64) | beta(64)] of the same 64 channels and the kernel writes
//     leaky_relu(((x - mean) * rstd) * (1 + gamma) + beta, 0.2) -- gamma, beta and the normalised tensor never reach HBM.
//   * K chunks are ordered [ky][64-channel chunk][kx].  For the stride-1 modes (all but four layers of the network) the
//     three kx taps of a group are the SAME pixels shifted by one operand row, so a gather thread loads its pixel once per
//     group and stores it into the three operand tiles at rows r+1, r, r-1 (zero rows at the image's left / right edge;
//     the two pixels beyond the tile's ends come from a two-lane halo warp): a third of the global loads of a
//     tap-by-tap gather, which is what bounded the kernel (scripts/gpu_spade_conv_bound.py).
//   * warp roles: warp 0 weight producer, warp 1 MMA issuer, warps 2-5 gather (128 threads = 128 operand rows; the
//     loads of the next group are issued before the current one is stored), warp 6 halo, warps 7-10 epilogue.  The
//     accumulator is double buffered in TMEM (2 x 128 columns): the epilogue of tile t runs under the main loop of tile
//     t+1, and the gather runs straight through tile boundaries.  Bounded mbarrier waits (trap + status word), as
//     everywhere in this library.
#include <cuda_fp16.h>
#include "sahs_common.cuh"

namespace {

constexpr int kCThreads = 352;                       // warps: 0 weights, 1 MMA, 2-5 gather, 6 halo, 7-10 epilogue
constexpr int kCSlots = 3;
constexpr int kCChunk = 16384;                       // one [128 x 64] fp16 operand tile
constexpr int kCOffB = kCSlots * kCChunk;            // B ring after the A ring
constexpr int kCOffBar = 2 * kCSlots * kCChunk;      // 96 KB
constexpr int kCOffStage = kCOffBar + 256;           // floats: bias[128] | mean[64] | rstd[64]
constexpr int kCSmem = kCOffStage + 256 * 4;
constexpr int kCTmemCols = 256;                      // two accumulators of 128 columns

enum { MODE_S1 = 0, MODE_S2 = 1, MODE_T2 = 2, MODE_FIRST = 3, MODE_T2C = 4 };
enum { EPI_RELU = 1, EPI_ADD = 2, EPI_SPADE = 4, EPI_F32 = 8 };

struct ConvP {
  const __half* in;
  int in_h, in_w, in_cs, cin;
  int out_h, out_w;
  int mode, up, down;
  const uint8_t* packed_w;
  const float* bias;
  int ntile, ntiles_n, nchunks;
  int epi;
  const __half* aux;
  int aux_cs, aux_shift;
  const float* mean;
  const float* rstd;
  void* out;
  int out_cs, cout;
  int nkx, nky;          // taps per row / column of the kernel window (3 x 3; a parity class of a transposed conv: 1 or 2)
  int dy[2], dx[2];      // MODE_T2C: source offset of tap index ky / kx
  int rm, rm_w, rm_py, rm_px;   // MODE_T2C: tile pixel (i, j) is output pixel (2 i + py, 2 j + px) of a rm_w wide image
  int* status;
  int dbg;   // measurement switches (SAHS_CONV_DBG): 1 no gather loads, 2 no operand stores, 4 no MMAs, 8 16-byte weight copies, 16 no output stores, 32 tap-by-tap gather for the stride-1 modes too, 64 no zero-tap skipping in transposed convs
};

// fp32 pair -> fp16 pair, saturating at +-65504 like every other fp16 conversion of this library: an activation that
// outgrows fp16 is clamped instead of becoming inf (which InstanceNorm would turn into a frame of NaNs)
__device__ __forceinline__ __half2 sat_half2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return *reinterpret_cast<__half2*>(&r);
}

// input pixel of output pixel (oy, ox) for tap (ky, kx); false: outside (contributes zero)
__device__ __forceinline__ bool src_pixel(const ConvP& c, int oy, int ox, int ky, int kx, int& iy, int& ix) {
  if (c.mode == MODE_S2) {
    iy = 2 * oy + ky - 1;
    ix = 2 * ox + kx - 1;
    return iy >= 0 && iy < c.in_h && ix >= 0 && ix < c.in_w;
  }
  if (c.mode == MODE_T2C) {        // ky / kx: tap INDEX within the parity class; every tap is a plain shifted read
    iy = oy + c.dy[ky];
    ix = ox + c.dx[kx];
    return iy < c.in_h && ix < c.in_w;
  }
  if (c.mode == MODE_T2) {
    const int ty = oy + 1 - ky, tx = ox + 1 - kx;
    iy = ty >> 1;
    ix = tx >> 1;
    return ty >= 0 && tx >= 0 && !(ty & 1) && !(tx & 1) && iy < c.in_h && ix < c.in_w;
  }
  const int ly = oy + ky - 1, lx = ox + kx - 1;          // logical (resized) input has the output's size
  iy = (ly << c.down) >> c.up;
  ix = (lx << c.down) >> c.up;
  return ly >= 0 && ly < c.out_h && lx >= 0 && lx < c.out_w;
}

__global__ void __launch_bounds__(kCThreads, 2) spade_conv_kernel(const ConvP c) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kCOffBar);
  uint64_t* a_full = bars;                     // [3] 128 arrivals (one gather group)
  uint64_t* a_empty = bars + kCSlots;          // [3] tcgen05.commit
  uint64_t* b_full = bars + 2 * kCSlots;       // [3] expect_tx
  uint64_t* b_empty = bars + 3 * kCSlots;      // [3] tcgen05.commit
  uint64_t* acc_full = bars + 4 * kCSlots;     // [2] tcgen05.commit
  uint64_t* acc_empty = bars + 4 * kCSlots + 2;  // [2] 128 epilogue threads
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 4 * kCSlots + 4);
  float* stage = reinterpret_cast<float*>(smem + kCOffStage);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int nt = blockIdx.y;                                   // N tile of this CTA
  const long long P = (long long)c.out_h * c.out_w;
  const int tiles = (int)((P + 127) / 128);
  const bool fast = c.mode == MODE_S1 && !(c.dbg & 32);      // one load per (ky, chunk) feeds the three kx taps

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) { c.status[0] = 2; __trap(); }
    for (int i = 0; i < kCSlots; ++i) {
      mbar_init(&a_full[i], (fast && i != 1) ? 129 : 128);      // stride-1 path: + the halo lane of taps kx = 0 / 2
      mbar_init(&a_empty[i], 1);
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 128);
    }
    fence_mbar_init();
  }
  if (threadIdx.x < c.ntile) stage[threadIdx.x] = c.bias[nt * c.ntile + threadIdx.x];
  if ((c.epi & EPI_SPADE) && threadIdx.x < 64) {
    stage[128 + threadIdx.x] = c.mean[nt * 64 + threadIdx.x];
    stage[192 + threadIdx.x] = c.rstd[nt * 64 + threadIdx.x];
  }
  if (warp == 1) tmem_alloc(tmem_ptr, kCTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t b_bytes = (uint32_t)c.ntile * 128u;
  // Transposed conv: output row oy only receives the taps with (oy + 1 - ky) even.  When a tile lies inside one output
  // row (always, for widths that are multiples of 128) the chunks of the other ky are all zeros and every role skips them.
  const int kcs_all = c.cin >> 6;
  auto tile_row = [&](int tile) -> int {
    if (c.mode != MODE_T2 || (c.dbg & 64)) return -1;
    const long long p0 = (long long)tile * 128;
    if (p0 + 127 >= P) return -1;
    const int r0 = (int)(p0 / c.out_w), r1 = (int)((p0 + 127) / c.out_w);
    return r0 == r1 ? r0 : -1;
  };
  auto skipped = [&](int oyrow, int q) -> bool { return oyrow >= 0 && (((oyrow + 1 - (q / 3) / kcs_all) & 1) != 0); };

  if (warp == 0) {
    // ================= weights of this N tile: one [ntile x 64] block per K chunk, re-streamed per pixel tile (L2) ========
    uint32_t slot = 0, phase = 0;
    const uint8_t* wsrc = c.packed_w + (size_t)nt * c.nchunks * b_bytes;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      const int trow = tile_row(tile);
      for (int q = 0; q < c.nchunks; ++q) {
        if (skipped(trow, q)) continue;
        mbar_wait(&b_empty[slot], phase ^ 1, c.status, 100);
        if (lane == 0) {
          const uint32_t nbytes = (c.dbg & 8) ? 16u : b_bytes;
          mbar_arrive_expect_tx(&b_full[slot], nbytes);
          tma_bulk_g2s(smem + kCOffB + slot * kCChunk, wsrc + (size_t)q * b_bytes, nbytes, &b_full[slot]);
        }
        __syncwarp();
        if (++slot == kCSlots) { slot = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (converged warp, elect.sync per instruction) =================
    uint32_t slot = 0, phase = 0;
    const uint32_t idesc = umma_idesc_m128((uint32_t)c.ntile, true);
    int tcount = 0;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++tcount) {
      const int buf = tcount & 1;
      // accumulator `buf` was last used by tile tcount - 2: wait until its epilogue has read it (fresh barrier: passes)
      mbar_wait_uniform<false>(&acc_empty[buf], (uint32_t)(((tcount >> 1) & 1) ^ 1), c.status, 300);
      tc_fence_after();
      const uint32_t d = tmem_base + (uint32_t)buf * 128u;
      const int trow = tile_row(tile);
      uint32_t accumulate = 0u;
      for (int q = 0; q < c.nchunks; ++q) {
        if (skipped(trow, q)) continue;
        mbar_wait_uniform<false>(&a_full[slot], phase, c.status, 400);
        mbar_wait_uniform<false>(&b_full[slot], phase, c.status, 401);
        tc_fence_after();
        const uint64_t a0 = umma_smem_desc_sw128(smem_u32(smem + slot * kCChunk));
        const uint64_t b0 = umma_smem_desc_sw128(smem_u32(smem + kCOffB + slot * kCChunk));
        if (!(c.dbg & 4)) {
          tc_mma_f16_w(d, a0, b0, idesc, accumulate);
          tc_mma_f16_w(d, a0 + 2, b0 + 2, idesc, 1u);
          tc_mma_f16_w(d, a0 + 4, b0 + 4, idesc, 1u);
          tc_mma_f16_w(d, a0 + 6, b0 + 6, idesc, 1u);
        }
        accumulate = 1u;
        tc_commit_w(&a_empty[slot]);
        tc_commit_w(&b_empty[slot]);
        if (++slot == kCSlots) { slot = 0; phase ^= 1; }
      }
      tc_commit_w(&acc_full[buf]);
    }
  } else if (warp < 6) {
    // ================= gather: the A operand, one thread per row, next loads in flight =================
    // (a coalesced mapping -- eight lanes per pixel -- was measured too: the same time on the large layers, slower on
    // the small ones through its eight coordinate computations per chunk)
    const int row = (warp - 2) * 32 + lane;
    const int kcs = c.cin >> 6;                            // 64-channel chunks per tap (MODE_FIRST: unused)
    bool live = false;
    int oy = 0, ox = 0;
    auto set_tile = [&](int tile) {
      const long long p = (long long)tile * 128 + row;
      live = p < P;
      oy = live ? (int)(p / c.out_w) : 0;
      ox = live ? (int)(p - (long long)oy * c.out_w) : 0;
    };
    auto row_ptr = [&](int slot, int r) -> uint8_t* { return smem + slot * kCChunk + (r >> 3) * 1024 + (r & 7) * 128; };
    auto store_row = [&](uint8_t* dst, int r, const uint4 (&v)[8]) {
#pragma unroll
      for (int u = 0; u < 8; ++u) *reinterpret_cast<uint4*>(dst + (((u ^ r) & 7) << 4)) = v[u];
    };
    // chunk q = (ky * kcs + kc) * 3 + kx; group g = q / 3.  load(): the pixel of tap (ky, kx) for my row.
    auto load = [&](int g, int kx, uint4 (&v)[8]) {
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = make_uint4(0u, 0u, 0u, 0u);
      if (c.dbg & 1) return;
      if (c.mode == MODE_FIRST) {
        // nine taps x 4 channels (8 bytes per pixel) in one chunk: column = tap * 4 + channel
        uint2 t[10];
#pragma unroll
        for (int k = 0; k < 9; ++k) {
          int iy, ix;
          t[k] = make_uint2(0u, 0u);
          if (live && src_pixel(c, oy, ox, k / 3, k % 3, iy, ix))
            t[k] = __ldg(reinterpret_cast<const uint2*>(c.in + ((size_t)iy * c.in_w + ix) * c.in_cs));
        }
        t[9] = make_uint2(0u, 0u);
#pragma unroll
        for (int u = 0; u < 5; ++u) v[u] = make_uint4(t[2 * u].x, t[2 * u].y, t[2 * u + 1].x, t[2 * u + 1].y);
      } else {
        const int ky = g / kcs, kc = g - ky * kcs;
        int iy, ix;
        if (live && src_pixel(c, oy, ox, ky, kx, iy, ix)) {
          const uint4* src = reinterpret_cast<const uint4*>(c.in + ((size_t)iy * c.in_w + ix) * c.in_cs + kc * 64);
#pragma unroll
          for (int u = 0; u < 8; ++u) v[u] = __ldg(src + u);
        }
      }
    };
    int tile = blockIdx.x;
    uint4 vn[8];
    if (fast) {
      // ---- stride-1 modes: one load per group, three operand tiles (slot = kx) ----
      const int ngroups = 3 * kcs;
      uint32_t phase = 0;
      int g = 0;
      if (tile < tiles) {
        set_tile(tile);
        load(0, 1, vn);
      }
      while (tile < tiles) {
        uint4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = vn[u];
        const bool my_live = live;
        const int my_ox = ox;
        int ng = g + 1, ntile = tile;
        if (ng == ngroups) {
          ng = 0;
          ntile = tile + (int)gridDim.x;
          if (ntile < tiles) set_tile(ntile);
        }
        if (ntile < tiles) load(ng, 1, vn);                // requested before this group's slots are even free
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          mbar_wait(&a_empty[kx], phase ^ 1, c.status, 200 + kx);
          if (!(c.dbg & 2)) {
            // my pixel is tap kx of the output pixel one row up / the same / one row down (same image row only)
            const int dr = row - kx + 1, dx = my_ox - kx + 1;
            if (kx == 1 || (my_live && dr >= 0 && dr < 128 && dx >= 0 && dx < c.out_w)) store_row(row_ptr(kx, dr), dr, v);
            // my own row's tap kx lies outside the image (or I am beyond the last pixel): zeros, nobody else writes it
            if (kx != 1 && (!my_live || my_ox + kx - 1 < 0 || my_ox + kx - 1 >= c.out_w)) {
              const uint4 z[8] = {};
              store_row(row_ptr(kx, row), row, z);
            }
            fence_proxy_async_smem();
          }
          mbar_arrive(&a_full[kx]);
        }
        phase ^= 1;
        g = ng;
        tile = ntile;
      }
    } else {
      // ---- stride-2 / transposed / first-layer modes: tap by tap ----
      uint32_t slot = 0, phase = 0;
      int q = -1, trow = -1;
      auto advance = [&](int& t, int& qq, int& tr) {      // next chunk that is not skipped (set_tile on a tile change)
        do {
          if (qq < 0) { tr = tile_row(t); set_tile(t); }
          if (++qq == c.nchunks) {
            qq = 0;
            t += (int)gridDim.x;
            if (t >= tiles) return;
            tr = tile_row(t);
            set_tile(t);
          }
        } while (skipped(tr, qq));
      };
      if (tile < tiles) {
        advance(tile, q, trow);
        if (tile < tiles) load(q / c.nkx, q % c.nkx, vn);
      }
      while (tile < tiles) {
        uint4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = vn[u];
        int nq = q, ntile = tile, ntrow = trow;
        advance(ntile, nq, ntrow);
        if (ntile < tiles) load(nq / c.nkx, nq % c.nkx, vn);
        mbar_wait(&a_empty[slot], phase ^ 1, c.status, 210);
        if (!(c.dbg & 2)) {
          store_row(row_ptr((int)slot, row), row, v);
          fence_proxy_async_smem();
        }
        mbar_arrive(&a_full[slot]);
        if (++slot == kCSlots) { slot = 0; phase ^= 1; }
        q = nq;
        tile = ntile;
        trow = ntrow;
      }
    }
  } else if (warp == 6) {
    // ================= halo (stride-1 modes): the pixel left of the tile's first row (tap kx = 0) and right of its last
    // row (kx = 2), when they lie in the same image row =================
    if (fast && lane < 2) {
      const int kx = lane == 0 ? 0 : 2, drow = lane == 0 ? 0 : 127;
      const int kcs = c.cin >> 6, ngroups = 3 * kcs;
      bool need = false;
      int oy = 0, ox = 0;
      auto set_tile = [&](int tile) {
        const long long p = (long long)tile * 128 + drow;
        const bool live = p < P;
        oy = live ? (int)(p / c.out_w) : 0;
        ox = live ? (int)(p - (long long)oy * c.out_w) : 0;
        need = live && (kx == 0 ? ox >= 1 : ox + 1 < c.out_w);
      };
      auto load = [&](int g, uint4 (&v)[8]) {
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = make_uint4(0u, 0u, 0u, 0u);
        const int ky = g / kcs, kc = g - ky * kcs;
        int iy, ix;
        if (need && !(c.dbg & 1) && src_pixel(c, oy, ox, ky, kx, iy, ix)) {
          const uint4* src = reinterpret_cast<const uint4*>(c.in + ((size_t)iy * c.in_w + ix) * c.in_cs + kc * 64);
#pragma unroll
          for (int u = 0; u < 8; ++u) v[u] = __ldg(src + u);
        }
      };
      uint32_t phase = 0;
      int tile = blockIdx.x, g = 0;
      uint4 vn[8];
      if (tile < tiles) {
        set_tile(tile);
        load(0, vn);
      }
      while (tile < tiles) {
        uint4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = vn[u];
        const bool my_need = need;
        int ng = g + 1, ntile = tile;
        if (ng == ngroups) {
          ng = 0;
          ntile = tile + (int)gridDim.x;
          if (ntile < tiles) set_tile(ntile);
        }
        if (ntile < tiles) load(ng, vn);
        mbar_wait(&a_empty[kx], phase ^ 1, c.status, 220 + kx);
        if (my_need && !(c.dbg & 2)) {
          uint8_t* dst = smem + kx * kCChunk + (drow >> 3) * 1024 + (drow & 7) * 128;
#pragma unroll
          for (int u = 0; u < 8; ++u) *reinterpret_cast<uint4*>(dst + (((u ^ drow) & 7) << 4)) = v[u];
          fence_proxy_async_smem();
        }
        mbar_arrive(&a_full[kx]);
        phase ^= 1;
        g = ng;
        tile = ntile;
      }
    }
  } else {
    // ================= epilogue: TMEM -> registers -> NHWC, under the next tile's main loop =================
    const int quarter = warp & 3;                          // TMEM lane quarter this warp may read
    const int row = quarter * 32 + lane;                   // pixel row of the tile == TMEM lane
    int tcount = 0;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++tcount) {
      const int buf = tcount & 1;
      const long long p = (long long)tile * 128 + row;
      const bool live = p < P && !(c.dbg & 16);
      const int oy = live ? (int)(p / c.out_w) : 0, ox = live ? (int)(p - (long long)oy * c.out_w) : 0;
      mbar_wait(&acc_full[buf], (uint32_t)((tcount >> 1) & 1), c.status, 500);
      tc_fence_after();
      const uint32_t tmem_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)buf * 128u;
      const long long po = c.rm ? ((long long)(2 * oy + c.rm_py) * c.rm_w + 2 * ox + c.rm_px) : p;   // output pixel
      if (c.epi & EPI_SPADE) {
        // columns [0,64) gamma, [64,128) beta of channels nt*64 .. nt*64+63
        const long long ap = c.aux_shift ? ((long long)(oy >> c.aux_shift) * (c.out_w >> c.aux_shift) + (ox >> c.aux_shift)) : p;
        const __half* xrow = c.aux + (size_t)ap * c.aux_cs + nt * 64;
        __half* orow = reinterpret_cast<__half*>(c.out) + (size_t)p * c.out_cs + nt * 64;
#pragma unroll 1
        for (int i = 0; i < 64; i += 16) {
          uint32_t gm[16], bt[16];
          tmem_ld16(tmem_row + i, gm);
          tmem_ld16(tmem_row + 64 + i, bt);
          tmem_ld_wait();
          if (live) {
            uint4 xv[2];
            xv[0] = __ldg(reinterpret_cast<const uint4*>(xrow + i));
            xv[1] = __ldg(reinterpret_cast<const uint4*>(xrow + i) + 1);
            const __half* xh = reinterpret_cast<const __half*>(xv);
            uint4 ov[2];
            __half2* oh = reinterpret_cast<__half2*>(ov);
#pragma unroll
            for (int e = 0; e < 16; e += 2) {
              float r[2];
#pragma unroll
              for (int k = 0; k < 2; ++k) {
                const int ch = i + e + k;
                const float xn = (__half2float(xh[e + k]) - stage[128 + ch]) * stage[192 + ch];
                const float g = __uint_as_float(gm[e + k]) + stage[ch];
                const float b = __uint_as_float(bt[e + k]) + stage[64 + ch];
                const float y = xn * (1.f + g) + b;
                r[k] = y > 0.f ? y : 0.2f * y;
              }
              oh[e >> 1] = sat_half2(r[0], r[1]);
            }
            reinterpret_cast<uint4*>(orow + i)[0] = ov[0];
            reinterpret_cast<uint4*>(orow + i)[1] = ov[1];
          }
        }
      } else {
#pragma unroll 1
        for (int i = 0; i < c.ntile; i += 16) {
          uint32_t acc[16];
          tmem_ld16(tmem_row + i, acc);
          tmem_ld_wait();
          if (live) {
            const int ch0 = nt * c.ntile + i;
            float r[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              r[e] = __uint_as_float(acc[e]) + stage[i + e];
              if (c.epi & EPI_RELU) r[e] = fmaxf(r[e], 0.f);
            }
            if ((c.epi & EPI_ADD) && ch0 + 16 <= c.cout) {
              const uint4* ar = reinterpret_cast<const uint4*>(c.aux + (size_t)p * c.aux_cs + ch0);
              uint4 av[2] = {__ldg(ar), __ldg(ar + 1)};
              const __half* ah = reinterpret_cast<const __half*>(av);
#pragma unroll
              for (int e = 0; e < 16; ++e) r[e] += __half2float(ah[e]);
            }
            if (c.epi & EPI_F32) {
              float* orow = reinterpret_cast<float*>(c.out) + (size_t)po * c.out_cs;
#pragma unroll
              for (int e = 0; e < 16; ++e)
                if (ch0 + e < c.cout) orow[ch0 + e] = r[e];
            } else if (ch0 + 16 <= c.cout) {
              uint4 ov[2];
              __half2* oh = reinterpret_cast<__half2*>(ov);
#pragma unroll
              for (int e = 0; e < 16; e += 2) oh[e >> 1] = sat_half2(r[e], r[e + 1]);
              uint4* orow = reinterpret_cast<uint4*>(reinterpret_cast<__half*>(c.out) + (size_t)po * c.out_cs + ch0);
              orow[0] = ov[0];
              orow[1] = ov[1];
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&acc_empty[buf]);
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kCTmemCols);
  }
}

// ---- instance-norm statistics of an NHWC fp16 tensor: per-channel sum and sum of squares ----
// Deterministic (run-to-run bit-identical, so a CUDA-graph replay equals the eager forward): fp32 per thread over a
// fixed pixel sequence, fixed-order fp64 reduction over the block's pixel lanes, one fp64 partial per block, and a
// fixed-order sum over blocks in the finalize kernel.  No atomics.
constexpr int kStatBlocks = 256;      // upper bound of the grid; the workspace holds kStatBlocks x 2C doubles
__global__ void __launch_bounds__(256) instnorm_sums_kernel(const __half* __restrict__ x, long long P, int C, int cs,
                                                           double* __restrict__ partial) {
  __shared__ float sh[256 * 16];                     // [pixel lane][2C]
  const int groups = C >> 3;                         // 8-channel groups (uint4)
  const int cgp = threadIdx.x % groups, pl = threadIdx.x / groups, lanes = 256 / groups;
  float s[8], ss[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) s[e] = ss[e] = 0.f;
  for (long long p = (long long)blockIdx.x * lanes + pl; p < P; p += (long long)gridDim.x * lanes) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(x + (size_t)p * cs + cgp * 8));
    const __half* h = reinterpret_cast<const __half*>(&v);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float f = __half2float(h[e]);
      s[e] += f;
      ss[e] += f * f;
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    sh[pl * 2 * C + cgp * 8 + e] = s[e];
    sh[pl * 2 * C + C + cgp * 8 + e] = ss[e];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += 256) {
    double acc = 0.0;
    for (int l = 0; l < lanes; ++l) acc += (double)sh[l * 2 * C + i];
    partial[(size_t)blockIdx.x * 2 * C + i] = acc;
  }
}

// mean and 1/sqrt(biased variance + eps), ref: nn.InstanceNorm2d(affine=False), nerf/_init_spade.py:118
__global__ void instnorm_finalize_kernel(const double* __restrict__ partial, int nblocks, int C, double inv_count, float eps,
                                         float* __restrict__ mean, float* __restrict__ rstd) {
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= C) return;
  double su = 0.0, sq = 0.0;
  for (int b = 0; b < nblocks; ++b) {
    su += partial[(size_t)b * 2 * C + ch];
    sq += partial[(size_t)b * 2 * C + C + ch];
  }
  const double m = su * inv_count;
  double var = sq * inv_count - m * m;
  if (var < 0.0) var = 0.0;
  mean[ch] = (float)m;
  rstd[ch] = (float)(1.0 / sqrt(var + (double)eps));
}

// nn.AvgPool2d(2, stride=2) on NHWC fp16 (fp32 sum, one rounding)
__global__ void __launch_bounds__(256) avgpool2_kernel(const __half* __restrict__ x, int H, int W, int C,
                                                      __half* __restrict__ y) {
  const int groups = C >> 3;
  const long long total = (long long)(H / 2) * (W / 2) * groups;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % groups);
    const long long op = i / groups;
    const int ox = (int)(op % (W / 2)), oy = (int)(op / (W / 2));
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(x + ((size_t)(2 * oy + (k >> 1)) * W + 2 * ox + (k & 1)) * C + g * 8));
      const __half* h = reinterpret_cast<const __half*>(&v);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] += __half2float(h[e]);
    }
    uint4 o;
    __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int e = 0; e < 8; e += 2) oh[e >> 1] = __floats2half2_rn(0.25f * acc[e], 0.25f * acc[e + 1]);
    *reinterpret_cast<uint4*>(y + (size_t)op * C + g * 8) = o;
  }
}

}  // namespace

extern "C" int sahs_spade_conv(const sahs_conv_desc* d, void* stream) {
  SAHS_CHECK_ARG(d && d->in && d->packed_w && d->bias && d->out, "null pointer");
  SAHS_CHECK_ARG(d->out_h > 0 && d->out_w > 0 && d->in_h > 0 && d->in_w > 0 && d->out_h < 32768 && d->out_w < 32768,
                 "bad geometry");
  SAHS_CHECK_ARG(d->ntile == 16 || d->ntile == 64 || d->ntile == 128, "ntile must be 16, 64 or 128");
  SAHS_CHECK_ARG(d->ntiles >= 1 && d->ntiles <= 64, "bad N tile count");
  SAHS_CHECK_ARG(d->mode >= MODE_S1 && d->mode <= MODE_T2C, "bad mode");
  const bool spade = (d->epilogue & EPI_SPADE) != 0;
  if (d->mode == MODE_FIRST) {
    SAHS_CHECK_ARG(d->in_cs == 4 && d->in_h == d->out_h && d->in_w == d->out_w, "first-layer mode: [H,W,4] fp16 input");
  } else {
    SAHS_CHECK_ARG(d->cin >= 64 && d->cin % 64 == 0 && d->in_cs % 8 == 0 && d->in_cs >= d->cin, "cin must be a multiple of 64");
    SAHS_CHECK_ARG(((uintptr_t)d->in & 15) == 0, "input must be 16-byte aligned");
  }
  if (d->mode == MODE_S1 || d->mode == MODE_FIRST) {
    SAHS_CHECK_ARG(d->up_shift >= 0 && d->down_shift >= 0 && d->up_shift < 8 && d->down_shift < 8 &&
                       ((d->in_h << d->up_shift) >> d->down_shift) == d->out_h &&
                       ((d->in_w << d->up_shift) >> d->down_shift) == d->out_w &&
                       (d->up_shift == 0 || d->down_shift == 0),
                   "nearest resize must be an exact power of two");
  } else if (d->mode == MODE_S2) {
    SAHS_CHECK_ARG(d->out_h == (d->in_h + 1) / 2 && d->out_w == (d->in_w + 1) / 2, "stride-2 geometry");
  } else {
    SAHS_CHECK_ARG(d->out_h == 2 * d->in_h && d->out_w == 2 * d->in_w, "transposed-conv geometry");
    SAHS_CHECK_ARG(d->mode != MODE_T2C || (d->t2_class >= 0 && d->t2_class < 4 && !d->aux && !spade), "parity class 0..3, no aux");
  }
  if (spade) {
    SAHS_CHECK_ARG(d->ntile == 128 && d->aux && d->mean && d->rstd && !(d->epilogue & (EPI_RELU | EPI_ADD | EPI_F32)),
                   "SPADE epilogue: [gamma|beta] tiles of 128 columns, aux = the tensor being normalised");
    SAHS_CHECK_ARG(d->cout == d->ntiles * 64 && d->out_cs % 8 == 0 && d->aux_cs % 8 == 0 && d->aux_shift >= 0 && d->aux_shift < 8,
                   "SPADE epilogue geometry");
    SAHS_CHECK_ARG(d->aux_shift == 0 || ((d->out_h >> d->aux_shift) << d->aux_shift) == d->out_h, "aux_shift");
  } else {
    SAHS_CHECK_ARG(d->cout >= 1 && d->cout <= d->ntiles * d->ntile, "cout");
    SAHS_CHECK_ARG((d->epilogue & EPI_F32) || (d->cout % 16 == 0 && d->out_cs % 8 == 0), "fp16 outputs: cout multiple of 16");
    SAHS_CHECK_ARG(!(d->epilogue & EPI_ADD) || (d->aux && d->aux_cs % 8 == 0 && d->aux_shift == 0), "residual add needs aux");
  }
  SAHS_CHECK_ARG(((uintptr_t)d->out & 15) == 0 && (!d->aux || ((uintptr_t)d->aux & 15) == 0) && ((uintptr_t)d->packed_w & 15) == 0,
                 "pointers must be 16-byte aligned");
  static bool attr = false;
  if (!attr) {
    SAHS_CUDA(cudaFuncSetAttribute(spade_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kCSmem));
    attr = true;
  }
  ConvP c;
  c.in = (const __half*)d->in;
  c.in_h = d->in_h; c.in_w = d->in_w; c.in_cs = d->in_cs; c.cin = d->cin;
  c.out_h = d->out_h; c.out_w = d->out_w;
  c.mode = d->mode; c.up = d->up_shift; c.down = d->down_shift;
  c.packed_w = (const uint8_t*)d->packed_w;
  c.bias = d->bias;
  c.ntile = d->ntile; c.ntiles_n = d->ntiles;
  c.nkx = c.nky = 3;
  c.dy[0] = c.dy[1] = c.dx[0] = c.dx[1] = 0;
  c.rm = 0; c.rm_w = 0; c.rm_py = c.rm_px = 0;
  if (d->mode == MODE_T2C) {
    // output (2i + py, 2j + px) = sum over the taps with (o + 1 - k) even of input ((o + 1 - k) / 2):
    // parity 0: k = 1 -> input i;  parity 1: k = 0 -> input i + 1, k = 2 -> input i   (tap order as packed: ascending k)
    const int py = d->t2_class >> 1, px = d->t2_class & 1;
    c.nky = py ? 2 : 1; c.nkx = px ? 2 : 1;
    c.dy[0] = py ? 1 : 0; c.dy[1] = 0;
    c.dx[0] = px ? 1 : 0; c.dx[1] = 0;
    c.rm = 1; c.rm_w = d->out_w; c.rm_py = py; c.rm_px = px;
    c.out_h = d->in_h; c.out_w = d->in_w;              // tiles run over the input grid
  }
  c.nchunks = d->mode == MODE_FIRST ? 1 : c.nky * c.nkx * (d->cin / 64);
  c.epi = d->epilogue;
  c.aux = (const __half*)d->aux; c.aux_cs = d->aux_cs; c.aux_shift = d->aux_shift;
  c.mean = d->mean; c.rstd = d->rstd;
  c.out = d->out; c.out_cs = d->out_cs; c.cout = d->cout;
  c.status = sahs_status_words(3);
  { const char* e = getenv("SAHS_CONV_DBG"); c.dbg = e ? atoi(e) : 0; }
  SAHS_CHECK_ARG(c.status, "status word allocation failed");
  const long long P = (long long)c.out_h * c.out_w;
  const int tiles = (int)((P + 127) / 128);
  int gx = (2 * sahs_num_sms() + d->ntiles - 1) / d->ntiles;
  if (gx > tiles) gx = tiles;
  if (gx < 1) gx = 1;
  spade_conv_kernel<<<dim3((unsigned)gx, (unsigned)d->ntiles), kCThreads, kCSmem, (cudaStream_t)stream>>>(c);
  SAHS_LAUNCH_CHECK();
  return SAHS_OK;
}

extern "C" int sahs_instnorm_stats(const void* x, int64_t num_pixels, int channels, int channel_stride, float eps,
                                   double* sums_workspace, float* mean, float* rstd, void* stream) {
  SAHS_CHECK_ARG(x && sums_workspace && mean && rstd, "null pointer");
  SAHS_CHECK_ARG(num_pixels > 0 && channels >= 8 && channels <= 256 && channels % 8 == 0 && (256 % (channels / 8)) == 0 &&
                     channel_stride % 8 == 0 && channel_stride >= channels && ((uintptr_t)x & 15) == 0,
                 "channels must be 8..256 (a power-of-two number of 8-channel groups), 16-byte aligned rows");
  cudaStream_t s = (cudaStream_t)stream;
  const int lanes = 256 / (channels / 8);
  long long blocks = (num_pixels + lanes * 16 - 1) / (lanes * 16);
  if (blocks > kStatBlocks) blocks = kStatBlocks;
  if (blocks < 1) blocks = 1;
  instnorm_sums_kernel<<<(unsigned)blocks, 256, 0, s>>>((const __half*)x, num_pixels, channels, channel_stride, sums_workspace);
  SAHS_LAUNCH_CHECK();
  instnorm_finalize_kernel<<<(channels + 127) / 128, 128, 0, s>>>(sums_workspace, (int)blocks, channels,
                                                                  1.0 / (double)num_pixels, eps, mean, rstd);
  SAHS_LAUNCH_CHECK();
  return SAHS_OK;
}

extern "C" int sahs_avgpool2(const void* x, int height, int width, int channels, void* y, void* stream) {
  SAHS_CHECK_ARG(x && y && height >= 2 && width >= 2 && height % 2 == 0 && width % 2 == 0 && channels % 8 == 0 &&
                     ((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0,
                 "avgpool2: even height / width, channels multiple of 8");
  const long long total = (long long)(height / 2) * (width / 2) * (channels / 8);
  long long blocks = (total + 255) / 256;
  const long long cap = 8LL * sahs_num_sms();
  if (blocks > cap) blocks = cap;
  avgpool2_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const __half*)x, height, width, channels, (__half*)y);
  SAHS_LAUNCH_CHECK();
  return SAHS_OK;
}
