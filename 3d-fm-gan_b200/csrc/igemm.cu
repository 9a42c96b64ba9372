// Implicit-GEMM convolution on tcgen05 / TMEM, fed by TMA (sm_100a).
//
// GEMM view:  D[M = B*OH*OW pixels, N = Cout] = sum_{tap, c} A[pixel shifted by tap, c] * W[tap][n][c]
//   A: NHWC bf16 activations.  One 4-D TMA box (64 channels x tile_w x tile_h x tile_b) per
//      (tap, channel chunk) lands as a K-major SWIZZLE_128B [128 x 64] operand tile; the conv's
//      zero padding is TMA out-of-bounds fill, the conv stride is the TMA element stride.
//   W: bf16 [ntaps][rows][cin], one 2-D TMA box (64 x BLOCK_N) per (tap, chunk), shared by the batch.
//   D: fp32 in TMEM, double buffered (2 x BLOCK_N columns) so the epilogue of tile i overlaps
//      the MMAs of tile i+1.
// Warp roles (192 threads): warp 0 TMA producer, warp 1 MMA issuer + TMEM owner, warps 2-5
// epilogue (TMEM lane quarter = warp_idx % 4).  Persistent CTAs, static round-robin tiles.
//
// Epilogue (per element; tables staged in smem per tile):
//   v = acc*T0 + T1 + noise*noise_w (+ residual);  v = v>0 ? v : v*T2;  rgb += v*T[4..6];  out = v*T3
// which covers demodulation, noise injection, bias, leaky-ReLU*sqrt2, the next layer's style
// modulation and the fused ToRGB 1x1 conv (stylegan2.py:258-262, 312, 371, 393), and for the
// encoders folded BatchNorm + ReLU/PReLU/LeakyReLU + residual add.
//
// Algorithmic FLOPs per launch: 2 * B*OH*OW * Cin * Cout * ntaps   (SURVEY.md 8d).
#include <stdlib.h>

#include "common.cuh"

namespace fm {

constexpr int IG_BM = 128;     // pixels per tile (UMMA M)
constexpr int IG_BK = 64;      // channels per k-step (128 B rows, SWIZZLE_128B)
constexpr int IG_TRACE_N = 64;     // trace slots per CTA (debug)
constexpr int IG_HP_MAXA = 8;      // halo-patch slots (barrier pairs reserved)
constexpr int IG_MAX_STAGES = 12;  // operand-ring barrier pairs reserved (generic ring: Cfg::STAGES; halo-patch weight ring: hp_stages)
constexpr int IG_TILEQ = 32;      // entries of the dynamic tile queue
constexpr int IG_TAB_ROWS = 512;  // staged table rows per tile (tile_b_eff * BLOCK_N <= 512)

struct IgemmParams {
  int OH, OW, B;
  int tw, th, tb;                  // tile dims, tw*th*tb == 128
  int tiles_x, tiles_y, tiles_b, tiles_n, num_tiles;
  int kchunks, ntaps, stride_x, stride_y, w_rows, Cout;
  int rows;                        // tw*th*tb (<= 128) valid rows of the A tile
  int patch;                       // 3x3/s1/p1 row-patch mode: one 130-pixel A load serves the 3 horizontal taps
  int ksplit, kper;                // split-K: CTAs per (m,n) tile and k-iterations per CTA
  float* ws;                       // split-K fp32 workspace [B*OH*OW][ws_cs]
  int ws_cs;
  int cluster;                     // CTAs per cluster (1, 2 or 4): weight tiles are TMA-multicast to the cluster
  int pair;                        // cta_group::2: the 2 CTAs of a cluster run ONE M=256 MMA, each loads half of the weight tile
  int num_super;                   // tiles_n * ksplit * ceil(m_tiles / cluster)
  int m_tiles;
  int upmode;                      // fused stride-2 transposed 3x3 conv: 4 output-parity accumulators share the A tiles
  int nbuf;                        // TMEM accumulator buffers (2, or 1 when 4 x BN x 2 columns do not fit)
  int prows;                       // output rows per tile in patch mode (R accumulators share each weight load)
  int patch_a_bytes, patch_stage_bytes, patch_stages, patch_b_bytes;   // patch_b_bytes: stride between the 3 weight tiles of a stage
  // halo-patch mode (stride-1 tap sets): tile = 8 x 16 pixels of one image; ONE (16+dy span) x (8+dx span)
  // input patch per channel chunk serves every tap as a shifted UMMA descriptor (group stride = patch row)
  int hp, hp_pw, hp_ph, hp_bytes, hp_dx0, hp_dy0, hp_stages, hp_na, hp_dist;
  int gen_sbytes, gen_stages;      // generic ring: bytes per (A tile + weight tile) stage and stage count
  int hp_bstride;                  // bytes between weight stages of the halo-patch ring (half a tile per CTA of a pair)
  // strided convs: the taps split into parity planes of the input (iy = s*oy + dy -> plane dy mod s, row (dy - plane)/s of
  // the subsampled grid); each (chunk, plane) is one patch load (TMA element stride s) shared by the plane's taps
  int hp_np, hp_sx, hp_sy;
  int8_t hp_pl_first[5];           // taps [first[pl], first[pl+1]) belong to plane pl (taps are sorted by plane)
  int16_t hp_pl_x[4], hp_pl_y[4];  // input-space offset of a plane's patch origin relative to (sx*x0, sy*y0)
  // several output phases of a stride-2 transposed conv in one launch: super tile st -> tile t = st / nph, phase = (st % nph + t) % nph
  // (a cluster takes whole tiles, i.e. all phases of a tile back to back, starting with a different phase per tile);
  // taps [ph_first[ph], ph_first[ph+1]) and the output offset (ph_oy0, ph_ox0) belong to the phase
  int nph, nsup1;
  int ph_whole;                    // a cluster takes whole tiles (all phases back to back, rotated start) -- enough tiles per cluster
  int8_t ph_first[5];
  int16_t ph_oy0[4], ph_ox0[4];
  int hpw;                         // halo-patch mode with ALL weight tiles of the (single) n-tile resident in smem
  int wres, wres_stages;           // generic mode with all (tap, chunk) weight tiles of the single n-tile resident: only A tiles stream
  // strip (a wres sub-mode; 3-channel stems stored as 16-byte pixels, stride 1 in x): ONE compact box of the packed image
  // per tile -- strip_rows input rows x 136 pixels x 16 bytes -- serves every tap through a no-swizzle descriptor whose K
  // core stride is one pixel (the overlapping 8-pixel windows are never materialised: 6 KB per tile instead of 48 KB)
  int strip, strip_rows, strip_sbytes, strip_y0, strip_k32;
  int lean;                        // epilogue: the launch qualifies for the lean path (see the epilogue)
  int8_t strip_trow[FM_MAX_TAPS];  // input row of each tap inside the box
  int16_t hp_aoff[FM_MAX_TAPS];    // per-tap start offset of the A descriptor inside the patch (16-byte units)
  int Bg, nslabs;                  // images per group, weight slabs per group
  const float* border_tab;
  long long* trace;                // debug (FM3D_TRACE=1): per-CTA event timestamps [grid][IG_TRACE_N]
  int out_cgroup, cg_shrink;
  int max_ctas;                    // host only: cap of the persistent grid (0 = every SM)
  // EPI_RESUP: per tile, the box of the low-resolution residual that the tile's bilinear taps touch is staged in smem by
  // TMA (two channel slabs of res_slab_ch channels, pixel pitch res_pitch bytes = 16 mod 128: conflict-free 16-byte reads)
  int res_bw, res_bh, res_pitch, res_box_bytes, res_slab_bytes, res_slot_bytes, res_off, res_nslab;   // slab stride = box bytes rounded to 128
  float* colsum;                   // plain epilogue only: colsum[b][o] += sum over the tile's pixels of the stored value (SE squeeze)
  int* tile_ctr;                   // dynamic tile schedule: [0] next super tile, [1] clusters done (NULL = static round robin)
  long long out_gstride;
  void* out;
  int out_H, out_W, out_cstride, out_y0, out_x0, out_ys, out_xs, out_nchw_f32;
  int out_PH, out_PW;              // physical pitch of the NHWC output (rows per image, pixels per row) >= out_H, out_W
  const float* tab;
  int tab_bstride;
  const float* noise;
  int noise_bstride;
  const float* noise_w;
  const __nv_bfloat16* residual;
  int res_ih, res_iw;              // EPI_RESUP: size of the low-resolution residual source
  float res_ry, res_rx;            //            source step per output pixel ((ih-1)/(out_H-1), align_corners=True)
  float* rgb;
  int8_t tap_dy[FM_MAX_TAPS];
  int8_t tap_dx[FM_MAX_TAPS];
  int8_t tap_widx[FM_MAX_TAPS];
};

constexpr int IG_EPI_WARPS = 8;                       // two warps per TMEM lane quarter
constexpr int IG_THREADS2 = 64 + 32 * IG_EPI_WARPS;   // producer + MMA + epilogue warps

// epilogue feature flags (template: dead paths cost nothing)
constexpr int EPI_RGB = 1, EPI_RES = 2, EPI_BTAB = 4, EPI_SPLIT = 8, EPI_IDENT = 16;   // IDENT: out = acc (tab == NULL)
constexpr int EPI_RESUP = 32;   // residual = bilinear (align_corners) upsampling of a low-resolution tensor, sampled in the epilogue

template <int BN> struct IgemmCfg {
  static constexpr int A_BYTES = IG_BM * IG_BK * 2;        // 16 KB
  static constexpr int B_BYTES = BN * IG_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  // row-patch mode carves the same ring into stages of (R 130-pixel input rows + 3 weight tiles)
  static constexpr int RING_BYTES = 200 * 1024;
  static constexpr int TAB_BYTES = IG_TAB_ROWS * 32;
  static constexpr int RGB_BYTES = IG_BM * 16;
  static constexpr int SMEM_BYTES = RING_BYTES + TAB_BYTES + RGB_BYTES + 1024 /*barriers, tile queue*/;
  static constexpr int TMEM_COLS = 512;   // 2 buffers x R rows x BN columns; one CTA per SM owns all of TMEM
};

// Up mode (stride-2 transposed 3x3 conv, stylegan2.py:276): with T[2y+py, 2x+px] = sum over the kernel
// taps of matching parity, the four output parities need only four shifted views of the input
// (dy,dx in {0,-1}^2).  Stage s loads view s once and the weight taps that multiply it:
//   s=0 (0,0): W00->P00 W01->P01 W10->P10 W11->P11 | s=1 (0,-1): W02->P00 W12->P10
//   s=2 (-1,0): W20->P00 W21->P01                  | s=3 (-1,-1): W22->P00
__constant__ int8_t c_up_nb[4] = {4, 2, 2, 1};
__constant__ int8_t c_up_w[4][4] = {{0, 1, 3, 4}, {2, 5, 0, 0}, {6, 7, 0, 0}, {8, 0, 0, 0}};
__constant__ int8_t c_up_acc[4][4] = {{0, 1, 2, 3}, {0, 2, 0, 0}, {0, 1, 0, 0}, {0, 0, 0, 0}};
__constant__ int8_t c_up_dy[4] = {0, 0, -1, -1};
__constant__ int8_t c_up_dx[4] = {0, -1, 0, -1};

// Reader side of the dynamic tile queue (see igemm_conv_kernel): wait until slot k % IG_TILEQ carries the generation of
// the cluster's k-th tile and return the tile index.  Not inlined: called once per tile per warp, and its temporaries
// stay out of the register allocation of the epilogue (which sits at the 168-register ceiling).
__device__ __noinline__ int tileq_wait(const uint32_t* q, int k) {
  const uint32_t gen = ((static_cast<uint32_t>(k) / IG_TILEQ) & 1u) + 1u;
  const volatile uint32_t* slotp = reinterpret_cast<const volatile uint32_t*>(q + (k & (IG_TILEQ - 1)));
  uint32_t w = *slotp;
  if ((w >> 24) != gen) {
    const long long t0 = clock64();
    uint32_t spins = 0;
    while (((w = *slotp) >> 24) != gen) {
      __nanosleep(20);
      if ((++spins & 1023u) == 0 && clock64() - t0 > 4000000000LL) {
        printf("fm3d: tile queue wait timeout (block %d thread %d, k %d)\n", blockIdx.x, threadIdx.x, k);
        __trap();
      }
    }
  }
  return static_cast<int>(w & 0xffffffu);
}

// PAIR: cta_group::2 instantiation (a kernel that contains 2-CTA tcgen05 instructions can only be launched with an
// even cluster size, so the single-CTA paths live in their own instantiation).
template <int BN, int EPI, bool PAIR>
__global__ void __launch_bounds__(IG_THREADS2, 1)
igemm_conv_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmR, const IgemmParams p_in) {
  using Cfg = IgemmCfg<BN>;
  // kPair as a compile-time constant inside the kernel
  const IgemmParams& p = p_in;
  constexpr bool kPair = PAIR;
  // SWIZZLE_128B operand tiles need 1024-byte alignment; declaring it keeps the shared address space
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* s_stage = smem;
  float4* s_tab = reinterpret_cast<float4*>(smem + Cfg::RING_BYTES);
  float4* s_rgb = reinterpret_cast<float4*>(smem + Cfg::RING_BYTES + Cfg::TAB_BYTES);
  // border-correction table [9][BN] of the current n-tile: upper half of the table region + the (unused) RGB
  // scratch = 10 KB, available when the epilogue tables are shared by the batch and no ToRGB is fused
  float* s_btab = reinterpret_cast<float*>(smem + Cfg::RING_BYTES + Cfg::TAB_BYTES / 2);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::RING_BYTES + Cfg::TAB_BYTES + Cfg::RGB_BYTES);
  uint64_t* full_bar = bars;                        // [STAGES]  TMA -> MMA
  uint64_t* empty_bar = bars + IG_MAX_STAGES;       // [STAGES]  MMA -> TMA
  uint64_t* tfull_bar = bars + 2 * IG_MAX_STAGES;   // [2]       MMA -> epilogue
  uint64_t* tempty_bar = bars + 2 * IG_MAX_STAGES + 2;  // [2]   epilogue -> MMA
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2 * IG_MAX_STAGES + 4);
  uint64_t* afull_bar = bars + 2 * IG_MAX_STAGES + 5;   // [IG_HP_MAXA] halo-patch slots: TMA -> MMA
  uint64_t* aempty_bar = afull_bar + IG_HP_MAXA;      // [IG_HP_MAXA] MMA -> TMA
  // dynamic tile schedule: the cluster leader's producer draws super-tile indices from a global counter and publishes
  // them, in order, into every CTA's queue.  An entry is ONE word, (generation << 24) | tile with generation =
  // (k / IG_TILEQ) % 2 + 1 for the cluster's k-th tile (a slot holds 0, the previous round's entry or this round's):
  // readers poll the slot with plain volatile loads until the generation matches (no fences: nothing else is communicated through it; acquire/release at cluster scope compile
  // to MEMBAR.GPU and an L1 invalidate per poll).  The producer runs at most IG_MAX_STAGES + 1 tiles ahead of the MMA
  // warp (one stage per tile at least) and the MMA warp at most 2 tiles ahead of the epilogue, so a slot is never
  // overwritten while a reader still needs it.
  uint32_t* s_tileq = reinterpret_cast<uint32_t*>(aempty_bar + IG_HP_MAXA);
  uint64_t* rfull_bar = reinterpret_cast<uint64_t*>(s_tileq + IG_TILEQ);      // [2] EPI_RESUP: residual box landed (TMA -> epilogue)
  uint64_t* rempty_bar = rfull_bar + 2;                                        // [2] epilogue -> TMA
  static_assert((2 * IG_MAX_STAGES + 5 + 2 * IG_HP_MAXA) * 8 + IG_TILEQ * 4 + 4 * 8 <= 1024, "barrier region");
  static_assert(IG_TILEQ >= IG_MAX_STAGES + 1 + 2 + 4 && (IG_TILEQ & (IG_TILEQ - 1)) == 0, "tile queue depth");

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    if ((smem_u32(smem) & 1023u) != 0) { printf("fm3d: dynamic smem base not 1024-aligned\n"); __trap(); }
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (EPI & EPI_RESUP) tma_prefetch_desc(&tmR);
    for (int i = 0; i < IG_MAX_STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], kPair ? 1 : p.cluster);   // multicast commit of every CTA (cluster) / of the leader (pair)
    }
    for (int i = 0; i < IG_HP_MAXA; ++i) {
      mbar_init(&afull_bar[i], 1);
      mbar_init(&aempty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], IG_EPI_WARPS * (kPair ? 2 : 1));   // one arrive per epilogue warp (of both CTAs of a pair)
    }
    for (int i = 0; i < IG_TILEQ; ++i) reinterpret_cast<volatile uint32_t*>(s_tileq)[i] = 0;
    for (int i = 0; i < 2; ++i) {
      mbar_init(&rfull_bar[i], 1);
      mbar_init(&rempty_bar[i], IG_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    // Every CTA manages its own TMEM with cta_group::1 instructions, also under a cta_group::2 MMA (which only needs
    // the same column range in both CTAs: each allocates all 512 columns, base 0).  The cta_group::2 alloc / dealloc
    // couple the allocator state of the two SMs of the pair; with other streams' single-CTA kernels landing on one SM of
    // a TPC between two pair kernels that left a later pair's peer CTA blocked in tcgen05.alloc forever (observed with
    // cuda-gdb: one CTA at the cluster barrier, its peer's warp 1 polling in the alloc).
    tmem_alloc(s_tmem, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  if (p.cluster > 1) cluster_sync_all();           // barriers of every CTA are initialised before any remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  // programmatic dependent launch: everything above overlapped the tail of the previous kernel in the stream; the
  // next kernel may start taking SMs as soon as every CTA of this grid got here
  pdl_wait();
  pdl_launch_dependents();
  const int crank = p.cluster > 1 ? static_cast<int>(cluster_ctarank()) : 0;
  const int cluster_id = blockIdx.x / p.cluster, num_clusters = gridDim.x / p.cluster;
  const uint16_t cmask = static_cast<uint16_t>((1u << p.cluster) - 1);

  // ---- tile schedule.  tile_at(k) = the k-th super tile of this cluster (>= num_super: no more work).
  // Static: round robin over the clusters of the grid.  Dynamic (tile_ctr != NULL): drawn from a global counter, so a
  // CTA that got its SM late (another stream's kernel was still on it) simply takes fewer tiles instead of delaying the
  // whole grid -- what lets small layers of other networks run on a few SMs next to this kernel.
  const bool dyn = p.tile_ctr != nullptr;
  const bool q_leader = dyn && warp == 0 && crank == 0;
  // leader producer only.  The first tile of every cluster is its static one (no round trip to the counter before the
  // first load); tile k >= 1 is num_clusters + (a draw from the counter).  A draw is issued (tile_draw) one tile before
  // it is published, so its latency hides behind a tile's worth of loads.
  int q_fetched = 1;            // entries known so far (entry 0 is implicit)
  bool q_ended = false, q_pending = false;
  int q_id = 0;                 // lane 0: the draw in flight
  auto tile_draw = [&]() {
    if (q_leader && !q_pending && !q_ended) {
      if (lane == 0) q_id = atomicAdd(p.tile_ctr, 1);
      q_pending = true;
    }
  };
  auto tile_at = [&](int k) -> int {
    if (dyn && k == 0) return cluster_id;      // the counter hands out num_clusters, num_clusters + 1, ...
    if (!dyn) {
      if (p.ph_whole) {
        // several phases per tile: a cluster takes WHOLE tiles (all phases back to back), so every cluster gets the same
        // mix of 4- / 2- / 1-tap work, and (see ig_phase) starts each tile with a different phase: the clusters are in
        // different phases at any time and the chip-wide L2 demand is the average of the phases, not the 1-tap peak
        const int kk = k / p.nph, j = k - kk * p.nph;
        const int t = cluster_id + kk * num_clusters;
        return t < p.nsup1 ? t * p.nph + j : p.num_super;
      }
      return cluster_id + k * num_clusters;
    }
    if (q_leader) {
      while (q_fetched <= k && !q_ended) {
        tile_draw();
        int id = __shfl_sync(0xffffffffu, q_id, 0) + num_clusters;
        q_pending = false;
        if (id >= p.num_super) { id = p.num_super; q_ended = true; }
        if (lane == 0) {
          const uint32_t slot = smem_u32(&s_tileq[q_fetched & (IG_TILEQ - 1)]);
          const uint32_t word = (static_cast<uint32_t>(((q_fetched / IG_TILEQ) & 1) + 1) << 24) | static_cast<uint32_t>(id);
          for (int r = 0; r < p.cluster; ++r) st_cluster_u32(mapa_rank(slot, r), word);
        }
        ++q_fetched;
        __syncwarp();
      }
      if (k >= q_fetched) return p.num_super;      // past the end marker
    }
    return tileq_wait(s_tileq, k);
  };

  // super tile -> (tile, phase) when a launch carries several phases
  auto ig_phase = [&](int st, int& t) -> int {
    if (p.nph <= 1) { t = st; return 0; }
    if (!p.ph_whole) {                       // phase-major: all tiles of phase 0, then phase 1, ...
      const int ph = st / p.nsup1;
      t = st - ph * p.nsup1;
      return ph;
    }
    t = st / p.nph;
    const int j = st - t * p.nph + t;
    return j % p.nph;
  };
  const int kiters = p.ntaps * p.kchunks;
  // debug trace: slot i of this CTA <- clock (one writer per slot)
  long long* trc = p.trace ? p.trace + static_cast<size_t>(blockIdx.x) * IG_TRACE_N : nullptr;
#define IG_TRACE(slot) do { if (trc && (slot) < IG_TRACE_N) trc[(slot)] = clock64(); } while (0)
  if (threadIdx.x == 0) IG_TRACE(0);

  if (warp == 0) {
    // ============================== TMA producer ==============================
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t tx_bytes = static_cast<uint32_t>(p.rows) * (IG_BK * 2) + Cfg::B_BYTES;
    // halo-patch mode: patch cursor (tile, chunk) runs hp_dist chunks ahead of the weight cursor, across tiles
    int pk = 0, pst = (p.hp || p.hpw) ? tile_at(0) : 0, pkc = 0, pslot = 0;      // patch cursor: pst = tile_at(pk)
    tile_draw();
    uint32_t pphase = 0;
    const uint32_t patch_tx = static_cast<uint32_t>(p.hp_pw) * p.hp_ph * (IG_BK * 2);
    const int nvc = p.kchunks * p.hp_np;          // virtual chunks: (channel chunk, parity plane)
    // tile coordinates of the patch cursor: decomposed once per tile, and (chunk, plane) are counters -- the integer
    // divisions used to run once per patch (ncu source view of a merged-phase up-conv: a third of the producer's samples
    // in MUFU.RCP / IABS chains, ~300 clk per chunk against 512 clk of MMA time for a 1-tap chunk)
    int pbx = 0, pby = 0, pbb = 0, pkr = 0, ppl = 0;
    auto hp_tile_coords = [&]() {
      if (pst >= p.num_super) return;
      int pstl;
      ig_phase(pst, pstl);                                // the patch depends on the tile position only, not on the phase
      int m = (pstl / p.tiles_n) * p.cluster + crank;   // ksplit == 1 in this mode
      pbx = m % p.tiles_x; m /= p.tiles_x;
      pby = m % p.tiles_y;
      pbb = m / p.tiles_y;
    };
    hp_tile_coords();
    auto hp_prefetch = [&]() {
      if (pst >= p.num_super) return;
      const int bx = pbx, by = pby, bb = pbb;
      mbar_wait(&aempty_bar[pslot], pphase ^ 1);
      const int cx = bx * 8 * p.hp_sx + p.hp_pl_x[ppl], cy = by * 16 * p.hp_sy + p.hp_pl_y[ppl];
      if (lane == 0) {
        if (kPair) {
          // both CTAs' patches complete on the leader's barrier; the leader announces the bytes of both
          if (crank == 0) mbar_arrive_expect_tx(&afull_bar[pslot], 2 * patch_tx);
          tma_load_4d_pair(s_stage + pslot * p.hp_bytes, &tmA, mapa_rank(smem_u32(&afull_bar[pslot]), 0), pkr * IG_BK, cx, cy, bb);
        } else {
          mbar_arrive_expect_tx(&afull_bar[pslot], patch_tx);
          tma_load_4d(s_stage + pslot * p.hp_bytes, &tmA, &afull_bar[pslot], pkr * IG_BK, cx, cy, bb);
        }
      }
      __syncwarp();
      if (++ppl == p.hp_np) { ppl = 0; ++pkr; }
      if (++pkc == nvc) { pkc = 0; pkr = 0; pst = tile_at(++pk); tile_draw(); hp_tile_coords(); }
      if (++pslot == p.hp_na) { pslot = 0; pphase ^= 1; }
    };
    if (p.hpw) {
      // weight-resident variant: every (chunk, tap) weight tile of the single n-tile is loaded ONCE per CTA; after
      // that the producer only streams input patches (a TMA load costs ~5-8 clk per 128-byte row, so for small
      // channel counts the per-tile weight re-loads -- 64 rows per tap -- were 3/4 of all rows)
      const uint32_t wb = kPair ? Cfg::B_BYTES / 2 : Cfg::B_BYTES;
      uint8_t* sw = s_stage + p.hp_na * p.hp_bytes;
      if (lane == 0) {
        if (!kPair || crank == 0) mbar_arrive_expect_tx(&full_bar[0], static_cast<uint32_t>(kiters) * Cfg::B_BYTES);
        for (int kc = 0; kc < p.kchunks; ++kc)
          for (int tap = 0; tap < p.ntaps; ++tap) {
            uint8_t* dst = sw + static_cast<size_t>(kc * p.ntaps + tap) * wb;
            if (kPair) tma_load_2d_pair(dst, &tmB, mapa_rank(smem_u32(&full_bar[0]), 0), kc * IG_BK, p.tap_widx[tap] * p.w_rows + crank * (BN / 2));
            else tma_load_2d(dst, &tmB, &full_bar[0], kc * IG_BK, p.tap_widx[tap] * p.w_rows);
          }
      }
      __syncwarp();
      while (pst < p.num_super) hp_prefetch();                    // blocks on the patch ring only
    } else {
    if (p.hp)
      for (int i = 0; i < p.hp_dist; ++i) hp_prefetch();
    const bool any_tile = p.wres && tile_at(0) < p.num_super;
    if (p.wres && lane == 0 && any_tile) {
      // resident weights: every (tap, chunk) tile once per CTA (a TMA load costs ~5-6 clk per 128-byte row, and the
      // 3- and 7-tap stem convs spent a third of their rows re-loading the same 64 weight rows per tap for every tile)
      mbar_arrive_expect_tx(&afull_bar[0], static_cast<uint32_t>(kiters) * Cfg::B_BYTES);
      for (int it = 0; it < kiters; ++it) {
        const int tap = it / p.kchunks, kc = it - tap * p.kchunks;
        tma_load_2d(s_stage + static_cast<size_t>(it) * Cfg::B_BYTES, &tmB, &afull_bar[0], kc * IG_BK, p.tap_widx[tap] * p.w_rows);
      }
    }
    __syncwarp();
    if (lane == 0) IG_TRACE(1);
    int ptile = 0;
    int rslot = 0;
    uint32_t rphase = 0;
    for (int tk = 0;; ++tk) {
      const int st = tile_at(tk);
      if (st >= p.num_super) break;
      tile_draw();                    // tile tk + 1: drawn now, published when the producer gets there (the consumers are a
                                      // ring of stages behind), so the counter's round trip hides behind this tile's loads
      int stl;
      const int ph = ig_phase(st, stl);
      const int nt = stl % p.tiles_n;
      int m = stl / p.tiles_n;
      const int ks = m % p.ksplit; m /= p.ksplit;
      m = m * p.cluster + crank;                  // CTAs of a cluster take adjacent m-tiles of the same n-tile
      const int bx = m % p.tiles_x; m /= p.tiles_x;
      const int by = m % p.tiles_y;
      const int bb = m / p.tiles_y;
      const int x0 = bx * p.tw * p.stride_x, y0 = by * p.th * p.stride_y, b0 = bb * p.tb, n0 = nt * BN;
      const int it0 = ks * p.kper, it1 = min(kiters, it0 + p.kper);
      const int wrow0 = (b0 / p.Bg) * p.nslabs;
      if (p.upmode) {
        for (int it = 0; it < 4 * p.kchunks; ++it) {
          const int kc = it >> 2, sft = it & 3;
          const int nb = c_up_nb[sft];
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (lane == 0) {
            uint8_t* sa = s_stage + stage * (Cfg::A_BYTES + 4 * Cfg::B_BYTES);
            uint8_t* sb = sa + Cfg::A_BYTES;
            mbar_arrive_expect_tx(&full_bar[stage], static_cast<uint32_t>(p.rows) * (IG_BK * 2) + nb * Cfg::B_BYTES);
            tma_load_4d(sa, &tmA, &full_bar[stage], kc * IG_BK, x0 + c_up_dx[sft], y0 + c_up_dy[sft], b0);
            for (int j = 0; j < nb; ++j)
              { if (p.cluster == 1) tma_load_2d(sb + j * Cfg::B_BYTES, &tmB, &full_bar[stage], kc * IG_BK, (wrow0 + c_up_w[sft][j]) * p.w_rows + n0); else if (crank == 0) tma_load_2d_mcast(sb + j * Cfg::B_BYTES, &tmB, &full_bar[stage], kc * IG_BK, (wrow0 + c_up_w[sft][j]) * p.w_rows + n0, cmask); }
          }
          __syncwarp();
          if (++stage == p.patch_stages) { stage = 0; phase ^= 1; }
        }
        continue;
      }
      if (p.patch) {
        // stage = (kernel row ky, channel chunk): R 130-pixel input rows + the 3 weight tiles of that kernel row
        const uint32_t ptx = static_cast<uint32_t>(p.prows) * 130u * (IG_BK * 2) + 3u * Cfg::B_BYTES;
        for (int it = 0; it < 3 * p.kchunks; ++it) {
          const int ky = it / p.kchunks;
          const int kc = it - ky * p.kchunks;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (lane == 0 && kPair) {
            // CTA pair: own input rows, half of each of the 3 weight tiles; everything completes on the leader's barrier
            uint8_t* sa = s_stage + stage * p.patch_stage_bytes;
            uint8_t* sb = sa + p.patch_a_bytes;
            const uint32_t lbar = mapa_rank(smem_u32(&full_bar[stage]), 0);
            if (crank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2u * static_cast<uint32_t>(p.prows) * 130u * (IG_BK * 2) + 3u * Cfg::B_BYTES);
            tma_load_4d_pair(sa, &tmA, lbar, kc * IG_BK, x0 - 1, by * p.prows + ky - 1, b0);
#pragma unroll
            for (int kx = 0; kx < 3; ++kx)
              tma_load_2d_pair(sb + kx * p.patch_b_bytes, &tmB, lbar, kc * IG_BK, (wrow0 + ky * 3 + kx) * p.w_rows + n0 + crank * (BN / 2));
          } else if (lane == 0) {
            uint8_t* sa = s_stage + stage * p.patch_stage_bytes;
            uint8_t* sb = sa + p.patch_a_bytes;
            mbar_arrive_expect_tx(&full_bar[stage], ptx);
            tma_load_4d(sa, &tmA, &full_bar[stage], kc * IG_BK, x0 - 1, by * p.prows + ky - 1, b0);
#pragma unroll
            for (int kx = 0; kx < 3; ++kx)
              { if (p.cluster == 1) tma_load_2d(sb + kx * Cfg::B_BYTES, &tmB, &full_bar[stage], kc * IG_BK, (wrow0 + ky * 3 + kx) * p.w_rows + n0); else if (crank == 0) tma_load_2d_mcast(sb + kx * Cfg::B_BYTES, &tmB, &full_bar[stage], kc * IG_BK, (wrow0 + ky * 3 + kx) * p.w_rows + n0, cmask); }
          }
          __syncwarp();
          if (++stage == p.patch_stages) { stage = 0; phase ^= 1; }
        }
        continue;
      }
      if (p.hp) {
        // chunk-major weight stages; the input patch of chunk c + hp_dist is requested when chunk c starts
        // (its slot was last read hp_na - hp_dist >= 2 chunks ago, so the wait below never blocks the weights)
        uint8_t* sb0 = s_stage + p.hp_na * p.hp_bytes;
        for (int vc = 0, kc = 0, pl = 0; vc < nvc; ++vc) {
          hp_prefetch();
          const int tp0 = p.nph > 1 ? p.ph_first[ph] : p.hp_pl_first[pl], tp1 = p.nph > 1 ? p.ph_first[ph + 1] : p.hp_pl_first[pl + 1];
          for (int tap = tp0; tap < tp1; ++tap) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            if (lane == 0) {
              if (kPair) {
                if (crank == 0) mbar_arrive_expect_tx(&full_bar[stage], Cfg::B_BYTES);       // two halves
                tma_load_2d_pair(sb0 + stage * p.hp_bstride, &tmB, mapa_rank(smem_u32(&full_bar[stage]), 0), kc * IG_BK,
                                 (wrow0 + p.tap_widx[tap]) * p.w_rows + n0 + crank * (BN / 2));
              } else {
                mbar_arrive_expect_tx(&full_bar[stage], Cfg::B_BYTES);
                tma_load_2d(sb0 + stage * p.hp_bstride, &tmB, &full_bar[stage], kc * IG_BK, (wrow0 + p.tap_widx[tap]) * p.w_rows + n0);
              }
            }
            __syncwarp();
            if (++stage == p.hp_stages) { stage = 0; phase ^= 1; }
          }
          if (++pl == p.hp_np) { pl = 0; ++kc; }
        }
        if (lane == 0) IG_TRACE(2 + 4 * ptile);      // producer: all loads of this tile issued
        ++ptile;
        continue;
      }
      if (p.strip) {
        // weights resident; a stage holds the tile's strip of the packed image
        uint8_t* sa0 = s_stage + static_cast<size_t>(kiters) * Cfg::B_BYTES;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (lane == 0) {
          mbar_arrive_expect_tx(&full_bar[stage], static_cast<uint32_t>(p.strip_rows) * (136 * 16));
          tma_load_4d(sa0 + stage * p.strip_sbytes, &tmA, &full_bar[stage], 0, x0, y0 + p.strip_y0, b0);
        }
        __syncwarp();
        if (++stage == p.wres_stages) { stage = 0; phase ^= 1; }
        continue;
      }
      if (p.wres) {
        // weights were loaded once (below the tile loop); a stage holds one A tile
        uint8_t* sa0 = s_stage + static_cast<size_t>(kiters) * Cfg::B_BYTES;
        for (int it = 0; it < kiters; ++it) {
          const int tap = it / p.kchunks;
          const int kc = it - tap * p.kchunks;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (lane == 0) {
            mbar_arrive_expect_tx(&full_bar[stage], static_cast<uint32_t>(p.rows) * (IG_BK * 2));
            tma_load_4d(sa0 + stage * Cfg::A_BYTES, &tmA, &full_bar[stage], kc * IG_BK, x0 + p.tap_dx[tap], y0 + p.tap_dy[tap], b0);
          }
          __syncwarp();
          if (++stage == p.wres_stages) { stage = 0; phase ^= 1; }
        }
        continue;
      }
      if (EPI & EPI_RESUP) {
        // the low-resolution pixels this tile's bilinear taps touch: one box per channel slab, requested before the
        // tile's operands (its slot was released by the epilogue of the tile before last, which also owned the TMEM
        // buffer this tile's MMAs wait for)
        mbar_wait(&rempty_bar[rslot], rphase ^ 1);
        if (lane == 0) {
          const int sx0 = min(static_cast<int>(p.res_rx * (bx * p.tw + p.out_x0)), p.res_iw - 1);
          const int sy0 = min(static_cast<int>(p.res_ry * (by * p.th + p.out_y0)), p.res_ih - 1);
          uint8_t* dst = s_stage + p.res_off + rslot * p.res_slot_bytes;
          mbar_arrive_expect_tx(&rfull_bar[rslot], static_cast<uint32_t>(p.res_nslab * p.res_box_bytes));
          for (int sl = 0; sl < p.res_nslab; ++sl)
            tma_load_4d(dst + sl * p.res_slab_bytes, &tmR, &rfull_bar[rslot], n0 + sl * (BN < 128 ? BN : 128), sx0, sy0, b0);
        }
        __syncwarp();
        if (++rslot == 2) { rslot = 0; rphase ^= 1; }
      }
      for (int it = it0; it < it1; ++it) {
        const int tap = it / p.kchunks;
        const int kc = it - tap * p.kchunks;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (lane == 0 && kPair) {
          uint8_t* sa = s_stage + stage * p.gen_sbytes;
          const uint32_t lbar = mapa_rank(smem_u32(&full_bar[stage]), 0);
          if (crank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * static_cast<uint32_t>(p.rows) * (IG_BK * 2) + Cfg::B_BYTES);
          tma_load_4d_pair(sa, &tmA, lbar, kc * IG_BK, x0 + p.tap_dx[tap], y0 + p.tap_dy[tap], b0);
          tma_load_2d_pair(sa + Cfg::A_BYTES, &tmB, lbar, kc * IG_BK, (wrow0 + p.tap_widx[tap]) * p.w_rows + n0 + crank * (BN / 2));
        } else if (lane == 0) {
          uint8_t* sa = s_stage + stage * p.gen_sbytes;
          uint8_t* sb = sa + Cfg::A_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
          tma_load_4d(sa, &tmA, &full_bar[stage], kc * IG_BK, x0 + p.tap_dx[tap], y0 + p.tap_dy[tap], b0);
          { if (p.cluster == 1) tma_load_2d(sb, &tmB, &full_bar[stage], kc * IG_BK, (wrow0 + p.tap_widx[tap]) * p.w_rows + n0); else if (crank == 0) tma_load_2d_mcast(sb, &tmB, &full_bar[stage], kc * IG_BK, (wrow0 + p.tap_widx[tap]) * p.w_rows + n0, cmask); }
        }
        __syncwarp();
        if (++stage == p.gen_stages) { stage = 0; phase ^= 1; }
      }
    }
    }   // !hpw
  } else if (warp == 1) {
    // ============================== MMA issuer ==============================
    // The whole warp walks the loop in uniform control flow (so ptxas keeps stage indices and descriptors in
    // uniform registers) and one elected lane issues; the four K=16 MMAs of a 64-channel stage go out of one
    // asm block.  At N <= 128 a K=16 MMA lasts only 32-64 tensor-pipe cycles: every instruction between two
    // UTCHMMAs is on the critical path of the kernel.
    const uint32_t idesc = umma_idesc_bf16(kPair ? 2 * IG_BM : IG_BM, BN);
    constexpr uint32_t dhi = umma_desc_hi_sw128(1024);
    int stage = 0;
    uint32_t phase = 0;
    int titer = 0;
    int aslot = 0;
    uint32_t aslot_phase = 0;
    const uint32_t ring = smem_u32(s_stage);
    // Between the last MMA of a tile and the first of the next one this single warp runs a chain of dependent scalar
    // instructions while the tensor pipe idles (per-CTA traces: ~1100 clk per tile, a third of a 64-channel layer's
    // tile): everything loop-invariant is hoisted, and the accumulator buffer / phase are counters, not divisions.
    int buf = 0;
    uint32_t aphase = 0;
    auto next_acc = [&]() { if (++buf == p.nbuf) { buf = 0; aphase ^= 1; } };
    const uint32_t hp_wb = kPair ? Cfg::B_BYTES / 2 : Cfg::B_BYTES;
    const uint32_t hp_sw = ring + p.hp_na * p.hp_bytes;
    const uint32_t hp_ahi = umma_desc_hi_sw128(static_cast<uint32_t>(p.hp_pw) * 128u);
    const int hp_nvc = p.kchunks * p.hp_np;
    const bool hp_np1 = p.hp_np == 1;
    // CTA pair: only the leader issues (its MMAs read both CTAs' operands and write both CTAs' TMEM)
    for (;; ++titer, next_acc()) {
      const int st = (kPair && crank != 0) ? p.num_super : tile_at(titer);
      if (st >= p.num_super) break;
      mbar_wait(&tempty_bar[buf], aphase ^ 1);     // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + buf * BN;
      if (p.upmode) {
        const int nst = 4 * p.kchunks;
        const uint32_t tmem_t = tmem_base + buf * (4 * BN);
        const uint32_t sbytes = Cfg::A_BYTES + 4 * Cfg::B_BYTES;
        for (int it = 0; it < nst; ++it) {
          const int sft = it & 3;
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = ring + stage * sbytes;
          const int nb = c_up_nb[sft];
          if (elect_one()) {
            for (int j = 0; j < nb; ++j)
              umma_bf16_x4(tmem_t + c_up_acc[sft][j] * BN, umma_desc_lo(sa), dhi, umma_desc_lo(sa + Cfg::A_BYTES + j * Cfg::B_BYTES),
                           dhi, idesc, it > 0 ? 1u : 0u);
            if (p.cluster == 1) umma_commit(&empty_bar[stage]); else umma_commit_mcast(&empty_bar[stage], cmask);
            if (it == nst - 1) umma_commit(&tfull_bar[buf]);
          }
          __syncwarp();
          if (++stage == p.patch_stages) { stage = 0; phase ^= 1; }
        }
        continue;
      }
      if (p.patch) {
        const int nst = 3 * p.kchunks;
        const uint32_t tmem_t = tmem_base + buf * (p.prows * BN);
        for (int it = 0; it < nst; ++it) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = ring + stage * p.patch_stage_bytes;
          const uint32_t blo = umma_desc_lo(sa + p.patch_a_bytes);
          if (elect_one()) {
            for (int r = 0; r < p.prows; ++r) {
              // the tap's A operand is staged row r shifted by kx pixels (= kx 128-byte rows); the
              // 128B swizzle is a function of the smem address, so a shifted start address just works
              const uint32_t alo = umma_desc_lo(sa + r * (130 * 128));
              const uint32_t td = tmem_t + r * BN;
              if (kPair) {
                umma_bf16_x4_pair(td, alo, dhi, blo, dhi, idesc, it > 0 ? 1u : 0u);
                umma_bf16_x4_pair(td, alo + (128 >> 4), dhi, blo + (p.patch_b_bytes >> 4), dhi, idesc, 1u);
                umma_bf16_x4_pair(td, alo + (256 >> 4), dhi, blo + 2 * (p.patch_b_bytes >> 4), dhi, idesc, 1u);
              } else {
                umma_bf16_x4(td, alo, dhi, blo, dhi, idesc, it > 0 ? 1u : 0u);
                umma_bf16_x4(td, alo + (128 >> 4), dhi, blo + (Cfg::B_BYTES >> 4), dhi, idesc, 1u);
                umma_bf16_x4(td, alo + (256 >> 4), dhi, blo + 2 * (Cfg::B_BYTES >> 4), dhi, idesc, 1u);
              }
            }
            if (kPair) {
              umma_commit_pair(&empty_bar[stage]);
              if (it == nst - 1) umma_commit_pair(&tfull_bar[buf]);
            } else {
              if (p.cluster == 1) umma_commit(&empty_bar[stage]); else umma_commit_mcast(&empty_bar[stage], cmask);
              if (it == nst - 1) umma_commit(&tfull_bar[buf]);
            }
          }
          __syncwarp();
          if (++stage == p.patch_stages) { stage = 0; phase ^= 1; }
        }
        continue;
      }
      if (p.hpw) {
        const uint32_t wb = hp_wb, sw = hp_sw, ahi = hp_ahi;
        if (titer == 0) { mbar_wait(&full_bar[0], 0); tc_fence_after(); }      // resident weights have landed
        if (lane == 0) IG_TRACE(3 + 4 * titer);
        const int nvc = hp_nvc;
        for (int vc = 0, kc = 0, pl = 0; vc < nvc; ++vc) {
          mbar_wait(&afull_bar[aslot], aslot_phase);
          tc_fence_after();
          if (lane == 0 && vc == 0) IG_TRACE(4 + 4 * titer);
          const uint32_t alo0 = umma_desc_lo(ring + aslot * p.hp_bytes);
          const uint32_t blo0 = umma_desc_lo(sw + static_cast<uint32_t>(kc * p.ntaps) * wb);
          if (elect_one()) {
            for (int tap = p.hp_pl_first[pl]; tap < p.hp_pl_first[pl + 1]; ++tap) {
              const uint32_t alo = alo0 + static_cast<uint32_t>(p.hp_aoff[tap]);
              const uint32_t blo = blo0 + static_cast<uint32_t>(tap) * (wb >> 4);
              const uint32_t accf = (vc > 0 || tap > p.hp_pl_first[pl]) ? 1u : 0u;
              if (kPair) umma_bf16_x4_pair(tmem_d, alo, ahi, blo, dhi, idesc, accf);
              else umma_bf16_x4(tmem_d, alo, ahi, blo, dhi, idesc, accf);
            }
            if (kPair) {
              umma_commit_pair(&aempty_bar[aslot]);
              if (vc == nvc - 1) umma_commit_pair(&tfull_bar[buf]);
            } else {
              umma_commit(&aempty_bar[aslot]);
              if (vc == nvc - 1) umma_commit(&tfull_bar[buf]);
            }
          }
          __syncwarp();
          if (++aslot == p.hp_na) { aslot = 0; aslot_phase ^= 1; }
          if (++pl == p.hp_np) { pl = 0; ++kc; }
        }
        continue;
      }
      if (p.hp) {
        const uint32_t sb0 = hp_sw, ahi = hp_ahi;
        if (lane == 0) IG_TRACE(3 + 4 * titer);               // MMA: accumulator free
        const int nvc = hp_nvc;
        int mstl;
        const int mph = ig_phase(st, mstl);                   // once per tile (a division per chunk sat in the issue path)
        for (int vc = 0, pl = 0; vc < nvc; ++vc) {
          mbar_wait(&afull_bar[aslot], aslot_phase);          // this (chunk, plane)'s input patch has landed
          if (lane == 0 && vc == 0) IG_TRACE(4 + 4 * titer);  // MMA: first patch landed
          const uint32_t alo0 = umma_desc_lo(ring + aslot * p.hp_bytes);
          const int t0 = p.nph > 1 ? p.ph_first[mph] : p.hp_pl_first[pl], t1 = p.nph > 1 ? p.ph_first[mph + 1] : p.hp_pl_first[pl + 1];
          for (int tap = t0; tap < t1; ++tap) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t alo = alo0 + static_cast<uint32_t>(p.hp_aoff[tap]);
            const uint32_t blo = umma_desc_lo(sb0 + stage * p.hp_bstride);
            const uint32_t accf = (vc > 0 || tap > t0) ? 1u : 0u;
            if (elect_one()) {
              if (kPair) {
                umma_bf16_x4_pair(tmem_d, alo, ahi, blo, dhi, idesc, accf);
                umma_commit_pair(&empty_bar[stage]);
                if (tap == t1 - 1) {
                  umma_commit_pair(&aempty_bar[aslot]);
                  if (vc == nvc - 1) umma_commit_pair(&tfull_bar[buf]);
                }
              } else {
                umma_bf16_x4(tmem_d, alo, ahi, blo, dhi, idesc, accf);
                umma_commit(&empty_bar[stage]);
                if (tap == t1 - 1) {
                  umma_commit(&aempty_bar[aslot]);              // patch slot is free once these MMAs retire
                  if (vc == nvc - 1) umma_commit(&tfull_bar[buf]);
                }
              }
            }
            __syncwarp();
            if (++stage == p.hp_stages) { stage = 0; phase ^= 1; }
          }
          if (++aslot == p.hp_na) { aslot = 0; aslot_phase ^= 1; }
          if (++pl == p.hp_np) pl = 0;
        }
        continue;
      }
      if (p.strip) {
        if (titer == 0) { mbar_wait(&afull_bar[0], 0); tc_fence_after(); }      // resident weights have landed
        const uint32_t sa = ring + static_cast<uint32_t>(kiters) * Cfg::B_BYTES + stage * p.strip_sbytes;
        constexpr uint32_t ahi = umma_desc_hi_nosw(128);                         // 8 pixels of 16 bytes per row group
        if (lane == 0) IG_TRACE(3 + 4 * titer);                                  // MMA: accumulator free
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (lane == 0) IG_TRACE(4 + 4 * titer);                                  // MMA: strip landed
        if (elect_one()) {
          for (int tap = 0; tap < p.ntaps; ++tap) {
            const uint32_t alo = umma_desc_lo(sa + p.strip_trow[tap] * (136 * 16));
            const uint32_t blo = umma_desc_lo(ring + static_cast<uint32_t>(tap) * Cfg::B_BYTES);
            // Cin <= 32: the window's second half multiplies zero weights -- two K = 16 steps instead of four (traces: the
            // stem's tile period was the 12 no-swizzle MMAs of its three taps)
            if (p.strip_k32) umma_bf16_x2(tmem_d, alo, ahi, blo, dhi, idesc, tap > 0 ? 1u : 0u);
            else umma_bf16_x4(tmem_d, alo, ahi, blo, dhi, idesc, tap > 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          umma_commit(&tfull_bar[buf]);
        }
        __syncwarp();
        if (++stage == p.wres_stages) { stage = 0; phase ^= 1; }
        continue;
      }
      if (p.wres) {
        if (titer == 0) { mbar_wait(&afull_bar[0], 0); tc_fence_after(); }      // resident weights have landed
        const uint32_t sa0 = ring + static_cast<uint32_t>(kiters) * Cfg::B_BYTES;
        for (int it = 0; it < kiters; ++it) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t alo = umma_desc_lo(sa0 + stage * Cfg::A_BYTES), blo = umma_desc_lo(ring + static_cast<uint32_t>(it) * Cfg::B_BYTES);
          if (elect_one()) {
            umma_bf16_x4(tmem_d, alo, dhi, blo, dhi, idesc, it > 0 ? 1u : 0u);
            umma_commit(&empty_bar[stage]);
            if (it == kiters - 1) umma_commit(&tfull_bar[buf]);
          }
          __syncwarp();
          if (++stage == p.wres_stages) { stage = 0; phase ^= 1; }
        }
        continue;
      }
      const int ks = p.ksplit == 1 ? 0 : (st / p.tiles_n) % p.ksplit;
      const int it0 = ks * p.kper, it1 = min(kiters, it0 + p.kper);
      for (int it = it0; it < it1; ++it) {
        mbar_wait(&full_bar[stage], phase);        // TMA bytes have landed
        tc_fence_after();
        const uint32_t sa = ring + stage * p.gen_sbytes;
        const uint32_t alo = umma_desc_lo(sa), blo = umma_desc_lo(sa + Cfg::A_BYTES);
        if (elect_one()) {
          // +32 bytes per K=16 step inside the 128-byte swizzle row (address field is >>4)
          if (kPair) {
            umma_bf16_x4_pair(tmem_d, alo, dhi, blo, dhi, idesc, it > it0 ? 1u : 0u);
            umma_commit_pair(&empty_bar[stage]);
            if (it == it1 - 1) umma_commit_pair(&tfull_bar[buf]);
          } else {
            umma_bf16_x4(tmem_d, alo, dhi, blo, dhi, idesc, it > it0 ? 1u : 0u);
            if (p.cluster == 1) umma_commit(&empty_bar[stage]); else umma_commit_mcast(&empty_bar[stage], cmask);
            if (it == it1 - 1) umma_commit(&tfull_bar[buf]);
          }
        }
        __syncwarp();
        if (++stage == p.gen_stages) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ============================== epilogue (8 warps) ==============================
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;       // which half of the column chunks this warp takes
    const int row = q * 32 + lane;          // pixel index inside the tile
    const int etid = threadIdx.x - 64;      // 0..255
    const int tbe = p.tab_bstride ? p.tb : 1;
    const float nw = p.noise ? (p.noise_w ? __ldg(p.noise_w) : 1.f) : 0.f;
    const int lx = row % p.tw;
    const int ly = (row / p.tw) % p.th;
    const int lb = row / (p.tw * p.th);
    const float4* trow = s_tab + static_cast<size_t>(p.tab_bstride ? lb : 0) * BN * 2;
    const bool btab_smem = (EPI & EPI_BTAB) && !(EPI & EPI_RGB) && p.border_tab != nullptr && !p.tab_bstride &&
                           9 * BN * 4 <= Cfg::TAB_BYTES / 2 + Cfg::RGB_BYTES;
    int tab_key = -1;
    int titer = 0;
    const bool e_tn1 = p.tiles_n == 1, e_ty1 = p.tiles_y == 1, e_g1 = p.Bg == p.B;
    int buf = 0;
    uint32_t aphase = 0;
    // 32-byte stores when every 16-column chunk of a pixel starts on a 32-byte boundary
    const bool st256 = !p.out_nchw_f32 && (p.out_cstride & 15) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 31) == 0 &&
                       (p.out_gstride & 15) == 0;
    int eslot = 0;                 // EPI_RESUP: residual box slot / phase of the current tile
    uint32_t ephase = 0;
    // colsum: partial column sums [tile parity][lane quarter][BN] in the upper half of the table region (the tables
    // are shared by the batch there: BN rows of 32 bytes)
    float* s_csum = reinterpret_cast<float*>(smem + Cfg::RING_BYTES + Cfg::TAB_BYTES / 2);
    static_assert(2 * 4 * BN * 4 <= Cfg::TAB_BYTES / 2, "colsum scratch");
    auto next_acc = [&]() { if (++buf == p.nbuf) { buf = 0; aphase ^= 1; } };
    for (;; ++titer, next_acc()) {
      const int st = tile_at(titer);
      if (st >= p.num_super) break;
      // tile coordinates: the common single-n-tile / unsplit / ungrouped cases skip their integer divisions (each costs
      // ~25 dependent instructions, and a 64-channel tile's whole epilogue is only ~300 per warp)
      int stl;
      const int eph = ig_phase(st, stl);
      const int nt = e_tn1 ? 0 : stl % p.tiles_n;
      int m = e_tn1 ? stl : stl / p.tiles_n;
      if (p.ksplit > 1) m /= p.ksplit;
      m = m * p.cluster + crank;
      const int m2 = m / p.tiles_x;
      const int bx = m - m2 * p.tiles_x;
      const int bb = e_ty1 ? m2 : m2 / p.tiles_y;
      const int by = e_ty1 ? 0 : m2 - bb * p.tiles_y;
      const int n0 = nt * BN;
      const int ox = bx * p.tw + lx, b = bb * p.tb + lb;
      const int grp = e_g1 ? 0 : min((bb * p.tb) / p.Bg, p.B / p.Bg - 1);   // padded cluster tiles clamp to the last group

      // ---- epilogue tables: re-staged only when (sample block | group, n-tile) changes
      const int key = (p.tab_bstride ? bb : grp) * p.tiles_n + nt;
      if (!(EPI & (EPI_SPLIT | EPI_IDENT)) && key != tab_key) {
        tab_key = key;
        asm volatile("bar.sync 1, 256;" ::: "memory");      // everyone is done with the old tables
        for (int i = etid; i < tbe * BN; i += 256) {
          const int sb = i / BN, j = i - sb * BN;
          const int o = n0 + j;
          const int bsrc = p.tab_bstride ? (bb * p.tb + sb) : grp;
          float4 t0 = make_float4(0.f, 0.f, 0.f, 0.f), t1 = t0;
          if (o < p.Cout && (!p.tab_bstride || bsrc < p.B)) {
            const float4* src = reinterpret_cast<const float4*>(p.tab + (static_cast<size_t>(bsrc) * p.Cout + o) * 8);
            t0 = __ldg(src);
            t1 = __ldg(src + 1);
          }
          s_tab[2 * i] = t0;
          s_tab[2 * i + 1] = t1;
        }
        if (btab_smem) {
          for (int i = etid; i < 9 * BN; i += 256) {
            const int cls = i / BN, o = n0 + (i - cls * BN);
            s_btab[i] = o < p.Cout ? __ldg(p.border_tab + static_cast<size_t>(cls) * p.Cout + o) : 0.f;
          }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }


      // ---- lean path (host flag p.lean): plain or residual epilogue, bf16 NHWC output, every 16-column chunk inside
      // Cout, one accumulator per tile, no noise.  The general loop below spends ~310 instructions per tile and ~225 per
      // chunk and warp (ncu source view of a 64-channel ResNet layer), half of them address arithmetic, bounds tests and
      // constant-bank reloads for features these launches do not use -- and per-CTA traces of the 3-channel stem showed
      // the epilogue warps, not TMA or the tensor pipe, setting the tile period (3600 clk for 768 clk of MMAs).
      bool lean_done = false;
      if constexpr (EPI == 0 || EPI == EPI_RES) {
        if (p.lean) {
          lean_done = true;
          const int oy = by * p.th + ly;
          const bool valid = row < p.rows && ox < p.OW && oy < p.OH && b < p.B;
          const int Y = oy * p.out_ys + p.out_y0, X = ox * p.out_xs + p.out_x0;
          const uint32_t tmem_acc = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * BN;
          __nv_bfloat16* const op = static_cast<__nv_bfloat16*>(p.out) +
                                    ((static_cast<size_t>(b) * p.out_PH + Y) * p.out_PW + X) * p.out_cstride + n0;
          const __nv_bfloat16* const rp = (EPI & EPI_RES) ? p.residual + ((static_cast<size_t>(b) * p.out_H + Y) * p.out_W + X) * p.out_cstride + n0
                                                          : nullptr;
          const int nch = min(BN, p.Cout - n0) >> 4;           // chunks of this n-tile (Cout % 16 == 0)
          uint4 rn0 = make_uint4(0, 0, 0, 0), rn1 = rn0;
          if ((EPI & EPI_RES) && valid && half < nch) {
            rn0 = __ldg(reinterpret_cast<const uint4*>(rp + half * 16));
            rn1 = __ldg(reinterpret_cast<const uint4*>(rp + half * 16) + 1);
          }
          mbar_wait(&tfull_bar[buf], aphase);
          tc_fence_after();
          if (threadIdx.x == 64) IG_TRACE(5 + 4 * titer);
#pragma unroll 1
          for (int c = half; c < nch; c += 2) {
            uint32_t acc[16];
            tmem_ld_32x16(tmem_acc + c * 16, acc);
            uint32_t rw[8];
            if (EPI & EPI_RES) {
              rw[0] = rn0.x; rw[1] = rn0.y; rw[2] = rn0.z; rw[3] = rn0.w; rw[4] = rn1.x; rw[5] = rn1.y; rw[6] = rn1.z; rw[7] = rn1.w;
              if (valid && c + 2 < nch) {                        // the next chunk's residual, one iteration ahead
                rn0 = __ldg(reinterpret_cast<const uint4*>(rp + (c + 2) * 16));
                rn1 = __ldg(reinterpret_cast<const uint4*>(rp + (c + 2) * 16) + 1);
              }
            }
            const float4* tr = trow + 2 * (c * 16);
            tmem_ld_wait();
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float4 t0 = tr[2 * j];
              float x = fmaf(__uint_as_float(acc[j]), t0.x, t0.y);
              if (EPI & EPI_RES) x += __uint_as_float((j & 1) ? (rw[j >> 1] & 0xffff0000u) : (rw[j >> 1] << 16));
              x = x > 0.f ? x : x * t0.z;
              v[j] = x * t0.w;
            }
            if constexpr (EPI == 0) {
              if (p.colsum != nullptr) {
                float s8[8], s4[4], s2[2];
                const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4, h2 = lane & 2;
                if (!__all_sync(0xffffffffu, valid)) {
#pragma unroll
                  for (int j = 0; j < 16; ++j) v[j] = valid ? v[j] : 0.f;
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) s8[j] = (h16 ? v[8 + j] : v[j]) + __shfl_xor_sync(0xffffffffu, h16 ? v[j] : v[8 + j], 16);
#pragma unroll
                for (int j = 0; j < 4; ++j) s4[j] = (h8 ? s8[4 + j] : s8[j]) + __shfl_xor_sync(0xffffffffu, h8 ? s8[j] : s8[4 + j], 8);
#pragma unroll
                for (int j = 0; j < 2; ++j) s2[j] = (h4 ? s4[2 + j] : s4[j]) + __shfl_xor_sync(0xffffffffu, h4 ? s4[j] : s4[2 + j], 4);
                float s1 = (h2 ? s2[1] : s2[0]) + __shfl_xor_sync(0xffffffffu, h2 ? s2[0] : s2[1], 2);
                s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
                const int colj = (h16 ? 8 : 0) + (h8 ? 4 : 0) + (h4 ? 2 : 0) + (h2 ? 1 : 0);
                if (!(lane & 1)) s_csum[((titer & 1) * 4 + q) * BN + c * 16 + colj] = s1;
              }
            }
            if (valid) {
              uint4 w0, w1;
              w0.x = pack_bf16x2(v[0], v[1]); w0.y = pack_bf16x2(v[2], v[3]); w0.z = pack_bf16x2(v[4], v[5]); w0.w = pack_bf16x2(v[6], v[7]);
              w1.x = pack_bf16x2(v[8], v[9]); w1.y = pack_bf16x2(v[10], v[11]); w1.z = pack_bf16x2(v[12], v[13]); w1.w = pack_bf16x2(v[14], v[15]);
              uint4* o4 = reinterpret_cast<uint4*>(op + c * 16);
              if (st256) {
                st_global_256(o4, w0, w1);
              } else {
                o4[0] = w0;
                o4[1] = w1;
              }
            }
          }
        }
      }
      if (!lean_done) {
      const int nrows = p.upmode ? 4 : (p.patch ? p.prows : 1);
#pragma unroll 1
      for (int r = 0; r < nrows; ++r) {
      const int oy = p.patch ? by * p.prows + r : by * p.th + ly;
      // up mode: accumulator r holds output parity (r>>1, r&1); odd parities have one row/column less
      const int upy = p.upmode ? (r >> 1) : 0, upx = p.upmode ? (r & 1) : 0;
      const bool valid = row < p.rows && ox < p.OW - upx && oy < p.OH - upy && b < p.B;
      const uint32_t tmem_acc = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * (nrows * BN) + r * BN;
      const int Y = oy * p.out_ys + (p.nph > 1 ? p.ph_oy0[eph] : p.out_y0) + upy, X = ox * p.out_xs + (p.nph > 1 ? p.ph_ox0[eph] : p.out_x0) + upx;
      float nz = 0.f;
      if (valid && p.noise)
        nz = nw * __ldg(p.noise + (static_cast<size_t>(p.noise_bstride ? b : 0) * p.out_H + Y) * p.out_W + X);
      const size_t pix = (static_cast<size_t>(b) * p.out_H + Y) * p.out_W + X;        // logical: noise, residual, rgb
      const size_t opix = (static_cast<size_t>(b) * p.out_PH + Y) * p.out_PW + X;     // physical: the NHWC output
      float r0 = 0.f, r1 = 0.f, r2 = 0.f;
      // EPI_RESUP: the four source pixels and weights of this thread's output pixel (F.interpolate, bilinear,
      // align_corners=True: psp_encoders.py:81-98 -- the FPN top-down path, fused here instead of materialised)
      uint32_t ru00 = 0, ru01 = 0, ru10 = 0, ru11 = 0;       // byte offsets of the four taps inside the staged box
      float rw00 = 0.f, rw01 = 0.f, rw10 = 0.f, rw11 = 0.f;
      const uint8_t* rbox = nullptr;
      if (EPI & EPI_RESUP) {
        const float fy = p.res_ry * Y, fx = p.res_rx * X;
        const int y0 = min(static_cast<int>(fy), p.res_ih - 1), x0 = min(static_cast<int>(fx), p.res_iw - 1);
        const int y1 = min(y0 + 1, p.res_ih - 1), x1 = min(x0 + 1, p.res_iw - 1);
        const float ly = fy - y0, lx = fx - x0;
        // box origin = the taps of the tile's first pixel (the producer computes the same); rows / columns past the
        // tile (threads whose pixel is not valid) are clamped into the box and their results are never stored
        const int sx0 = min(static_cast<int>(p.res_rx * (bx * p.tw + p.out_x0)), p.res_iw - 1);
        const int sy0 = min(static_cast<int>(p.res_ry * (by * p.th + p.out_y0)), p.res_ih - 1);
        const int u0 = min(max(x0 - sx0, 0), p.res_bw - 1), u1 = min(max(x1 - sx0, 0), p.res_bw - 1);
        const int v0 = min(max(y0 - sy0, 0), p.res_bh - 1), v1 = min(max(y1 - sy0, 0), p.res_bh - 1);
        ru00 = static_cast<uint32_t>((v0 * p.res_bw + u0) * p.res_pitch); ru01 = static_cast<uint32_t>((v0 * p.res_bw + u1) * p.res_pitch);
        ru10 = static_cast<uint32_t>((v1 * p.res_bw + u0) * p.res_pitch); ru11 = static_cast<uint32_t>((v1 * p.res_bw + u1) * p.res_pitch);
        rw00 = (1.f - ly) * (1.f - lx); rw01 = (1.f - ly) * lx; rw10 = ly * (1.f - lx); rw11 = ly * lx;
        rbox = s_stage + p.res_off + eslot * p.res_slot_bytes;
      }
      const float* btab = nullptr;      // folded-input-BN border correction (per-thread: thread = pixel)
      bool warp_has_border = false;
      int cls = 0;
      if (EPI & EPI_BTAB) {
        cls = (oy == 0 ? 1 : (oy == p.OH - 1 ? 2 : 0)) * 3 + (ox == 0 ? 1 : (ox == p.OW - 1 ? 2 : 0));
        if (cls && p.border_tab) btab = p.border_tab + static_cast<size_t>(cls) * p.Cout;
        warp_has_border = __any_sync(0xffffffffu, btab != nullptr);   // interior warps skip the correction code
      }

      // residual of this thread's first chunk: requested before the accumulator wait, later chunks one iteration ahead
      uint4 resn[2];
      auto res_fetch = [&](int c) {
        const int o0 = n0 + c * 16;
        const int ol0 = o0 - (p.out_cgroup ? o0 / p.out_cgroup : 0) * p.out_cgroup;
        const uint4* rp = reinterpret_cast<const uint4*>(p.residual + pix * p.out_cstride + ol0);
#pragma unroll
        for (int g = 0; g < 2; ++g)
          resn[g] = (valid && p.residual && c < BN / 16 && ol0 + 8 * g < p.out_cstride) ? __ldg(rp + g) : make_uint4(0, 0, 0, 0);
      };
      if (EPI & EPI_RES) res_fetch(half);
      if (r == 0) {
        // everything above (noise, residual, bilinear taps) is in flight while the MMAs of this tile finish
        if (EPI & EPI_RESUP) mbar_wait(&rfull_bar[eslot], ephase);
        mbar_wait(&tfull_bar[buf], aphase);
        tc_fence_after();
        if (threadIdx.x == 64) IG_TRACE(5 + 4 * titer);         // epilogue: accumulator complete
      }

#pragma unroll 1
      for (int c = half; c < BN / 16; c += 2) {     // 16-column chunks, alternating between the two warps
        uint32_t acc[16];
        tmem_ld_32x16(tmem_acc + c * 16, acc);
        const int o0 = n0 + c * 16;
        if (EPI & EPI_SPLIT) {
          // split-K partial: raw fp32 accumulators are reduced into the workspace with vector atomics;
          // igemm_splitk_finalize_kernel applies the epilogue
          tmem_ld_wait();
          if (valid && o0 < p.Cout) {
            float* wp = p.ws + (static_cast<size_t>(b) * p.OH * p.OW + static_cast<size_t>(oy) * p.OW + ox) * p.ws_cs + o0;
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              if (o0 + j < p.Cout)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(wp + j), "f"(__uint_as_float(acc[j])),
                             "f"(__uint_as_float(acc[j + 1])), "f"(__uint_as_float(acc[j + 2])),
                             "f"(__uint_as_float(acc[j + 3]))
                             : "memory");
            }
          }
          continue;
        }
        // position inside the output tensor (concatenated-N outputs are written group-major)
        const int og = p.out_cgroup ? o0 / p.out_cgroup : 0;
        const int ol0 = o0 - og * p.out_cgroup;
        uint4 resv[2];
        if (EPI & EPI_RES) {
          resv[0] = resn[0]; resv[1] = resn[1];
          res_fetch(c + 2);
        }
        float resf[16];
        if (EPI & EPI_RESUP) {
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            // channels [c*16 + 8g, +8) of the n-tile: slab (c*16) / 128, 16 bytes at pitch res_pitch (16 mod 128: the 8
            // lanes of a shared-memory wavefront read different bank groups, lanes with the same source pixel broadcast)
            constexpr int SLAB_CH = BN < 128 ? BN : 128;
            const uint8_t* rb = rbox + ((c * 16) / SLAB_CH) * p.res_slab_bytes + ((c * 16) % SLAB_CH) * 2 + 16 * g;
            const uint4 qa = *reinterpret_cast<const uint4*>(rb + ru00), qb = *reinterpret_cast<const uint4*>(rb + ru01);
            const uint4 qc = *reinterpret_cast<const uint4*>(rb + ru10), qd = *reinterpret_cast<const uint4*>(rb + ru11);
            const uint32_t wa[4] = {qa.x, qa.y, qa.z, qa.w}, wb2[4] = {qb.x, qb.y, qb.z, qb.w};
            const uint32_t wc[4] = {qc.x, qc.y, qc.z, qc.w}, wd[4] = {qd.x, qd.y, qd.z, qd.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              resf[8 * g + 2 * k] = rw00 * __uint_as_float(wa[k] << 16) + rw01 * __uint_as_float(wb2[k] << 16) +
                                    rw10 * __uint_as_float(wc[k] << 16) + rw11 * __uint_as_float(wd[k] << 16);
              resf[8 * g + 2 * k + 1] = rw00 * __uint_as_float(wa[k] & 0xffff0000u) + rw01 * __uint_as_float(wb2[k] & 0xffff0000u) +
                                        rw10 * __uint_as_float(wc[k] & 0xffff0000u) + rw11 * __uint_as_float(wd[k] & 0xffff0000u);
            }
          }
        }
        float bcor[16];
        if (EPI & EPI_BTAB) {
          // border pixels (one warp in four on a 128-px row tile) fetch their 16 corrections as 4 vector loads
          if (warp_has_border && btab_smem) {
            // class 0 (interior) rows of the table are zero-valued for this purpose: read row `cls` only on the border
            const float4* bs = reinterpret_cast<const float4*>(s_btab + cls * BN + c * 16);
#pragma unroll
            for (int g4 = 0; g4 < 4; ++g4) {
              const float4 bv = btab != nullptr ? bs[g4] : make_float4(0.f, 0.f, 0.f, 0.f);
              bcor[4 * g4 + 0] = bv.x; bcor[4 * g4 + 1] = bv.y; bcor[4 * g4 + 2] = bv.z; bcor[4 * g4 + 3] = bv.w;
            }
          } else if (warp_has_border) {
#pragma unroll
            for (int g4 = 0; g4 < 4; ++g4) {
              float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
              if (btab != nullptr && (p.Cout & 3) == 0 && o0 + 4 * g4 + 3 < p.Cout) {
                bv = __ldg(reinterpret_cast<const float4*>(btab + o0 + 4 * g4));
              } else if (btab != nullptr) {
                if (o0 + 4 * g4 + 0 < p.Cout) bv.x = __ldg(btab + o0 + 4 * g4 + 0);
                if (o0 + 4 * g4 + 1 < p.Cout) bv.y = __ldg(btab + o0 + 4 * g4 + 1);
                if (o0 + 4 * g4 + 2 < p.Cout) bv.z = __ldg(btab + o0 + 4 * g4 + 2);
                if (o0 + 4 * g4 + 3 < p.Cout) bv.w = __ldg(btab + o0 + 4 * g4 + 3);
              }
              bcor[4 * g4 + 0] = bv.x; bcor[4 * g4 + 1] = bv.y; bcor[4 * g4 + 2] = bv.z; bcor[4 * g4 + 3] = bv.w;
            }
          }
        }
        tmem_ld_wait();
        // chunks past Cout only zero-fill the pad channels of an NHWC tensor (warp-uniform test)
        if (o0 >= p.Cout && (p.out_nchw_f32 || p.out_cgroup || ol0 >= p.out_cstride)) continue;
        const float4* tr = trow + 2 * (c * 16);
        float v[16];
        if (EPI & EPI_IDENT) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(acc[j]);
        } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float4 t0 = tr[2 * j];
          float x = fmaf(__uint_as_float(acc[j]), t0.x, t0.y + nz);
          if (EPI & EPI_BTAB) {
            if (warp_has_border) x += bcor[j];
          }
          if (EPI & EPI_RES) {
            const uint32_t w = (&resv[0].x)[j >> 1];
            x += __uint_as_float((j & 1) ? (w & 0xffff0000u) : (w << 16));
          }
          if (EPI & EPI_RESUP) x += resf[j];
          x = x > 0.f ? x : x * t0.z;
          if (EPI & EPI_RGB) {
            const float4 t1 = tr[2 * j + 1];
            r0 = fmaf(x, t1.x, r0);
            r1 = fmaf(x, t1.y, r1);
            r2 = fmaf(x, t1.z, r2);
          }
          v[j] = x * t0.w;
        }
        }
        if constexpr (EPI == 0) {
          if (p.colsum != nullptr) {
            // per-channel sums of the tile's outputs (the squeeze of an SE block, psp model_irse bottleneck_IR_SE: the
            // AdaptiveAvgPool2d(1) over this conv's output): 16 columns x 32 rows per warp, reduced by a transposing
            // butterfly (16 shuffles: each step halves the columns a lane carries) and added with one fp32 reduction
            // per column and warp.  tile_b == 1: the warp's rows belong to one image.
            float s8[8], s4[4], s2[2];
            const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4, h2 = lane & 2;
            if (!__all_sync(0xffffffffu, valid)) {       // partial tiles only: rows past the image do not count
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = valid ? v[j] : 0.f;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float lo = v[j], hi = v[8 + j];
              s8[j] = (h16 ? hi : lo) + __shfl_xor_sync(0xffffffffu, h16 ? lo : hi, 16);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) s4[j] = (h8 ? s8[4 + j] : s8[j]) + __shfl_xor_sync(0xffffffffu, h8 ? s8[j] : s8[4 + j], 8);
#pragma unroll
            for (int j = 0; j < 2; ++j) s2[j] = (h4 ? s4[2 + j] : s4[j]) + __shfl_xor_sync(0xffffffffu, h4 ? s4[j] : s4[2 + j], 4);
            float s1 = (h2 ? s2[1] : s2[0]) + __shfl_xor_sync(0xffffffffu, h2 ? s2[0] : s2[1], 2);
            s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
            const int colj = (h16 ? 8 : 0) + (h8 ? 4 : 0) + (h4 ? 2 : 0) + (h2 ? 1 : 0);      // the column this lane pair ended up with
            // the four lane quarters (and the rows of a row-patch tile) meet in shared memory -- each warp owns its slots
            // (quarter q, the columns of its chunks), so plain stores / adds: a shared-memory float atomicAdd is a CAS spin
            // loop -- and one thread per column then issues one global reduction per tile
            if (!(lane & 1)) {
              float* slot = s_csum + ((titer & 1) * 4 + q) * BN + c * 16 + colj;
              *slot = (r == 0 ? 0.f : *slot) + s1;
            }
          }
        }
        if (valid && p.out && ox < p.OW - og * p.cg_shrink) {   // out == NULL: only the fused ToRGB sums are wanted
          if (p.out_nchw_f32) {
            const size_t plane = static_cast<size_t>(p.out_H) * p.out_W;
            float* op = static_cast<float*>(p.out) + (static_cast<size_t>(b) * p.Cout + o0) * plane +
                        static_cast<size_t>(Y) * p.out_W + X;
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (o0 + j < p.Cout) op[j * plane] = v[j];
          } else {
            uint4* op = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) + og * p.out_gstride +
                                                 opix * p.out_cstride + ol0);
            uint4 w[2];
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              w[g].x = pack_bf16x2(v[8 * g + 0], v[8 * g + 1]);
              w[g].y = pack_bf16x2(v[8 * g + 2], v[8 * g + 3]);
              w[g].z = pack_bf16x2(v[8 * g + 4], v[8 * g + 5]);
              w[g].w = pack_bf16x2(v[8 * g + 6], v[8 * g + 7]);
            }
            if (st256 && ol0 + 16 <= p.out_cstride) {
              st_global_256(op, w[0], w[1]);                 // one whole 32-byte sector per lane
            } else {
#pragma unroll
              for (int g = 0; g < 2; ++g)
                if (ol0 + 8 * g < p.out_cstride) op[g] = w[g];
            }
          }
        }
      }
      if (EPI & EPI_RGB) {
        // combine the two column halves of each pixel through smem (deterministic), then write
        if (half == 1) s_rgb[row] = make_float4(r0, r1, r2, 0.f);
        asm volatile("bar.sync 2, 256;" ::: "memory");
        if (half == 0 && valid && p.rgb) {
          const float4 o = s_rgb[row];
          float* rp = p.rgb + pix * 4;
          if (p.tiles_n == 1) {
            *reinterpret_cast<float4*>(rp) = make_float4(r0 + o.x, r1 + o.y, r2 + o.z, 0.f);
          } else {
            atomicAdd(rp + 0, r0 + o.x);
            atomicAdd(rp + 1, r1 + o.y);
            atomicAdd(rp + 2, r2 + o.z);
          }
        }
        asm volatile("bar.sync 2, 256;" ::: "memory");   // s_rgb may be overwritten by the next tile
      }
      }   // rows of the tile
      }   // general path
      if constexpr (EPI == 0) {
        if (p.colsum != nullptr) {
          asm volatile("bar.sync 3, 256;" ::: "memory");         // every warp's partial sums of this tile are in place
          // (double buffered by tile parity: the next tile's stores go to the other half, and the barrier of the tile
          // after that orders them behind these reads)
          if (etid < BN && n0 + etid < p.Cout && bb * p.tb < p.B) {
            const float* sc = s_csum + (titer & 1) * 4 * BN + etid;
            atomicAdd(p.colsum + static_cast<size_t>(bb * p.tb) * p.Cout + n0 + etid, (sc[0] + sc[BN]) + (sc[2 * BN] + sc[3 * BN]));
          }
        }
      }
      // accumulator drained -> hand the TMEM buffer back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (EPI & EPI_RESUP) {
        if (lane == 0) mbar_arrive(&rempty_bar[eslot]);      // this warp no longer reads the residual box
        if (++eslot == 2) { eslot = 0; ephase ^= 1; }
      }
      if (lane == 0) {
        if (kPair) mbar_arrive_cluster_relaxed(mapa_rank(smem_u32(&tempty_bar[buf]), 0));   // the leader's MMA warp waits for both CTAs
        else mbar_arrive(&tempty_bar[buf]);
      }
      if (threadIdx.x == 64 && (p.hpw || p.strip)) IG_TRACE(2 + 4 * titer);    // epilogue (warp 2) drained this tile
    }
  }

  if (threadIdx.x == 64) IG_TRACE(IG_TRACE_N - 1);           // epilogue finished its last tile
  if (dyn && crank == 0 && threadIdx.x == 0) {
    // this cluster drew its last index (the end marker) earlier in this thread's program order; the last cluster to
    // get here re-arms both counters for the next launch that shares them (stream-ordered after this grid)
    __threadfence();
    if (atomicAdd(p.tile_ctr + 1, 1) == num_clusters - 1) {
      atomicExch(p.tile_ctr, 0);
      atomicExch(p.tile_ctr + 1, 0);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (p.cluster > 1) cluster_sync_all();           // no CTA exits while a peer may still multicast into it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// Split-K tail: workspace (fp32 sums) -> epilogue -> bf16 NHWC, and re-zero the workspace.
// All loads of a thread (2 workspace vectors, 8 table rows, residual, noise) are issued up front.
__global__ void __launch_bounds__(256) igemm_splitk_finalize_kernel(const IgemmParams p, int64_t total) {
  pdl_wait();
  pdl_launch_dependents();
  const int groups8 = p.ws_cs / 8;
  const float nw = p.noise ? (p.noise_w ? __ldg(p.noise_w) : 1.f) : 0.f;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    const int g = static_cast<int>(idx % groups8);
    const int64_t pixlin = idx / groups8;
    const int ox = static_cast<int>(pixlin % p.OW);
    const int oy = static_cast<int>((pixlin / p.OW) % p.OH);
    const int b = static_cast<int>(pixlin / (static_cast<int64_t>(p.OW) * p.OH));
    const int o0 = g * 8;
    float4* wp = reinterpret_cast<float4*>(p.ws + pixlin * p.ws_cs + o0);
    const float4 a0 = wp[0], a1 = wp[1];
    const int Y = oy * p.out_ys + p.out_y0, X = ox * p.out_xs + p.out_x0;
    const size_t pix = (static_cast<size_t>(b) * p.out_H + Y) * p.out_W + X;
    const size_t opix = (static_cast<size_t>(b) * p.out_PH + Y) * p.out_PW + X;
    const bool store = o0 < p.out_cstride;
    const int trow = p.tab_bstride ? b : b / p.Bg;
    const float4* tp = reinterpret_cast<const float4*>(p.tab + static_cast<size_t>(trow) * p.Cout * 8);
    float4 t[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) t[j] = __ldg(tp + 2 * min(o0 + j, p.Cout - 1));     // clamped: no branch, loads batch
    uint4 res = make_uint4(0, 0, 0, 0);
    if (p.residual && store) res = __ldg(reinterpret_cast<const uint4*>(p.residual + pix * p.out_cstride + o0));
    float nz = 0.f;
    if (p.noise) nz = nw * __ldg(p.noise + (static_cast<size_t>(p.noise_bstride ? b : 0) * p.out_H + Y) * p.out_W + X);
    wp[0] = make_float4(0.f, 0.f, 0.f, 0.f);
    wp[1] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!store) continue;
    const float r0 = __uint_as_float(res.x << 16), r1 = __uint_as_float(res.x & 0xffff0000u);
    const float r2 = __uint_as_float(res.y << 16), r3 = __uint_as_float(res.y & 0xffff0000u);
    const float r4 = __uint_as_float(res.z << 16), r5 = __uint_as_float(res.z & 0xffff0000u);
    const float r6 = __uint_as_float(res.w << 16), r7 = __uint_as_float(res.w & 0xffff0000u);
#define FM_FIN(J, ACC, RES)                                              \
    float v##J = fmaf(ACC, t[J].x, t[J].y + nz) + RES;                   \
    v##J = v##J > 0.f ? v##J : v##J * t[J].z;                            \
    v##J = (o0 + J < p.Cout) ? v##J * t[J].w : 0.f;
    FM_FIN(0, a0.x, r0) FM_FIN(1, a0.y, r1) FM_FIN(2, a0.z, r2) FM_FIN(3, a0.w, r3)
    FM_FIN(4, a1.x, r4) FM_FIN(5, a1.y, r5) FM_FIN(6, a1.z, r6) FM_FIN(7, a1.w, r7)
#undef FM_FIN
    uint4 w;
    w.x = pack_bf16x2(v0, v1); w.y = pack_bf16x2(v2, v3);
    w.z = pack_bf16x2(v4, v5); w.w = pack_bf16x2(v6, v7);
    *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) + opix * p.out_cstride + o0) = w;
  }
}

// ------------------------------------------------------------------------------------ host
static long long* g_trace = nullptr;

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}

template <int BN, int EPI, bool PAIR>
static int launch_igemm3(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmR, const IgemmParams& p, cudaStream_t st) {
  using Cfg = IgemmCfg<BN>;
  static SmemOptIn opt_in;        // per instantiation, per device
  FM_CUDA_OK(smem_opt_in(opt_in, igemm_conv_kernel<BN, EPI, PAIR>, Cfg::SMEM_BYTES));
  int sms = sm_count();
  if (p.max_ctas > 0 && p.max_ctas < sms) sms = p.max_ctas < p.cluster ? p.cluster : p.max_ctas;
  const int max_clusters = sms / p.cluster;
  const int nclusters = p.num_super < max_clusters ? p.num_super : max_clusters;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(nclusters * p.cluster));
  cfg.blockDim = dim3(IG_THREADS2);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = static_cast<unsigned>(p.cluster);
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (pdl_enabled() && p.max_ctas <= 0) ? 2 : 1;
  FM_CUDA_OK(cudaLaunchKernelEx(&cfg, igemm_conv_kernel<BN, EPI, PAIR>, tmA, tmB, tmR, p));
  count_launch();
  FM_LAUNCH_OK();
  return FM_OK;
}

template <int BN, int EPI>
static int launch_igemm2(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmR, const IgemmParams& p, cudaStream_t st) {
  // pair instantiations exist for the feature sets the N >= 128 layers use
  if constexpr (EPI == 0 || EPI == EPI_RGB || EPI == EPI_RES || EPI == EPI_BTAB || EPI == EPI_IDENT || EPI == EPI_RESUP) {
    if (p.pair) return launch_igemm3<BN, EPI, true>(tmA, tmB, tmR, p, st);
  }
  if (p.pair) { set_error("fm_conv_igemm: internal: no CTA-pair instantiation for block_n %d epilogue %d", BN, EPI); return FM_ERR_INVALID; }
  return launch_igemm3<BN, EPI, false>(tmA, tmB, tmR, p, st);
}

template <int BN>
static int launch_igemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmR, const IgemmParams& p, cudaStream_t st) {
  // one instantiation per epilogue feature set in use (generator: RGB; ResNet: RES; IR block: BTAB)
  if (p.ksplit > 1) {
    const int rc = launch_igemm2<BN, EPI_SPLIT>(tmA, tmB, tmR, p, st);
    if (rc != FM_OK) return rc;
    const int64_t total = static_cast<int64_t>(p.B) * p.OH * p.OW * (p.ws_cs / 8);
    int64_t blocks = (total + 255) / 256;
    const int64_t cap = p.max_ctas > 0 ? static_cast<int64_t>(p.max_ctas) * 8 : static_cast<int64_t>(sm_count()) * 32;
    if (blocks > cap) blocks = cap;
    FM_CUDA_OK(launch_pdl(igemm_splitk_finalize_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, st, p, total));
    count_launch();
    return FM_OK;
  }
  if (!p.tab) return launch_igemm2<BN, EPI_IDENT>(tmA, tmB, tmR, p, st);
  if (p.residual && p.res_ih > 0) {
    if (p.rgb || p.border_tab) { set_error("fm_conv_igemm: an upsampled residual excludes rgb / border_tab"); return FM_ERR_INVALID; }
    return launch_igemm2<BN, EPI_RESUP>(tmA, tmB, tmR, p, st);
  }
  const int epi = (p.rgb ? EPI_RGB : 0) | (p.residual ? EPI_RES : 0) | (p.border_tab ? EPI_BTAB : 0);
  switch (epi) {
    case 0: return launch_igemm2<BN, 0>(tmA, tmB, tmR, p, st);
    case EPI_RGB: return launch_igemm2<BN, EPI_RGB>(tmA, tmB, tmR, p, st);
    case EPI_RES: return launch_igemm2<BN, EPI_RES>(tmA, tmB, tmR, p, st);
    case EPI_BTAB: return launch_igemm2<BN, EPI_BTAB>(tmA, tmB, tmR, p, st);
    default: return launch_igemm2<BN, EPI_RGB | EPI_RES | EPI_BTAB>(tmA, tmB, tmR, p, st);
  }
}

}  // namespace fm

extern "C" int fm_conv_igemm(const fm_conv_desc* d, void* stream) {
  using namespace fm;
  FM_CHECK_ARG(d != nullptr, "fm_conv_igemm: null desc");
  FM_CHECK_ARG(d->x && d->w && (d->out || d->rgb), "fm_conv_igemm: null tensor");
  FM_CHECK_ARG(d->tab || (!d->rgb && !d->residual && !d->border_tab && !d->noise && d->ksplit <= 1),
               "fm_conv_igemm: tab == NULL (identity epilogue) excludes rgb / residual / border_tab / noise / split-K");
  FM_CHECK_ARG(d->out || (!d->out_nchw_f32 && !d->out_cgroup && !d->upmode), "fm_conv_igemm: out may be NULL only for a plain NHWC conv with a fused RGB output");
  FM_CHECK_ARG(d->B > 0 && d->H > 0 && d->W > 0 && d->Cin > 0 && d->Cout > 0, "fm_conv_igemm: bad sizes");
  FM_CHECK_ARG(d->ntaps >= 1 && d->ntaps <= FM_MAX_TAPS, "fm_conv_igemm: ntaps %d out of range", d->ntaps);
  const int sx = d->stride_x > 0 ? d->stride_x : d->stride;
  const int sy = d->stride_y > 0 ? d->stride_y : d->stride;
  FM_CHECK_ARG(sx >= 1 && sx <= 8 && sy >= 1 && sy <= 8, "fm_conv_igemm: strides must be in 1..8 (got %d, %d)", sx, sy);
  const int64_t pixs = d->x_pixstride > 0 ? d->x_pixstride : d->x_cstride;
  const int64_t rows_ = d->x_rowstride > 0 ? d->x_rowstride : pixs * d->W;
  const int64_t imgs = d->x_imgstride > 0 ? d->x_imgstride : rows_ * d->H;
  FM_CHECK_ARG(pixs % 8 == 0 && rows_ % 8 == 0 && imgs % 8 == 0, "fm_conv_igemm: input strides must be multiples of 8 elements");
  FM_CHECK_ARG(d->x_pixstride > 0 || d->x_cstride >= d->Cin, "fm_conv_igemm: x_cstride < Cin");
  FM_CHECK_ARG(d->w_cstride % 8 == 0 && d->w_cstride >= d->Cin, "fm_conv_igemm: w_cstride must be a multiple of 8 and >= Cin");
  FM_CHECK_ARG(d->w_rows >= d->Cout, "fm_conv_igemm: w_rows < Cout");
  FM_CHECK_ARG((reinterpret_cast<uintptr_t>(d->x) & 15) == 0 && (reinterpret_cast<uintptr_t>(d->w) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(d->out) & 15) == 0, "fm_conv_igemm: tensors must be 16-byte aligned");
  FM_CHECK_ARG(d->out_nchw_f32 || d->out_cgroup || (d->out_cstride % 8 == 0 && d->out_cstride >= d->Cout),
               "fm_conv_igemm: out_cstride must be a multiple of 8 and >= Cout");
  FM_CHECK_ARG(d->OH > 0 && d->OW > 0 && d->out_ys >= 1 && d->out_xs >= 1, "fm_conv_igemm: bad output grid");
  FM_CHECK_ARG((d->OH - 1) * d->out_ys + d->out_y0 < d->out_H && (d->OW - 1) * d->out_xs + d->out_x0 < d->out_W,
               "fm_conv_igemm: output grid does not fit the output tensor");
  FM_CHECK_ARG(!(d->residual && (d->out_nchw_f32 || d->out_cgroup)), "fm_conv_igemm: residual needs a plain NHWC output");
  FM_CHECK_ARG(!d->border_tab || (d->OH >= 2 && d->OW >= 2), "fm_conv_igemm: border_tab needs OH, OW >= 2");
  FM_CHECK_ARG(!d->out_cgroup || (!d->out_nchw_f32 && d->out_cgroup % 32 == 0 && d->Cout % d->out_cgroup == 0 &&
                                  d->out_cstride % 8 == 0 && d->out_cstride >= d->out_cgroup && d->out_gstride % 8 == 0),
               "fm_conv_igemm: bad grouped-output parameters");
  FM_CHECK_ARG(!d->colsum || (d->tab && !d->tab_bstride && !d->rgb && !d->residual && !d->border_tab && !d->out_nchw_f32 && !d->out_cgroup && !d->upmode &&
                              d->nphases <= 1 && d->ksplit <= 1 && d->groups <= 1),
               "fm_conv_igemm: colsum needs the plain epilogue (no rgb / residual / border_tab / grouped or fp32 output / split-K)");
  const int G = d->groups > 1 ? d->groups : 1;
  FM_CHECK_ARG(d->B % G == 0, "fm_conv_igemm: batch %d not divisible by groups %d", d->B, G);
  const int Bg = d->B / G;
  const int nph = d->nphases > 1 ? d->nphases : 1;
  int ph_tmin = d->ntaps;
  if (nph > 1) {
    FM_CHECK_ARG(nph <= 4 && sx == 1 && sy == 1 && G == 1 && !d->upmode && !d->rgb && !d->residual && !d->border_tab && !d->noise,
                 "fm_conv_igemm: nphases > 1 needs a stride-1 ungrouped conv without rgb / residual / border_tab / noise");
    int tot = 0;
    for (int i = 0; i < nph; ++i) {
      FM_CHECK_ARG(d->phase_ntaps[i] >= 1 && d->phase_out_y0[i] >= 0 && d->phase_out_x0[i] >= 0, "fm_conv_igemm: bad phase %d", i);
      FM_CHECK_ARG((d->OH - 1) * d->out_ys + d->phase_out_y0[i] < d->out_H && (d->OW - 1) * d->out_xs + d->phase_out_x0[i] < d->out_W,
                   "fm_conv_igemm: phase %d does not fit the output tensor", i);
      tot += d->phase_ntaps[i];
      if (d->phase_ntaps[i] < ph_tmin) ph_tmin = d->phase_ntaps[i];
    }
    FM_CHECK_ARG(tot == d->ntaps, "fm_conv_igemm: phase_ntaps sum to %d, ntaps is %d", tot, d->ntaps);
  }

  EncodeTiledFn encode = get_encode_fn();
  if (!encode) { set_error("fm_conv_igemm: cuTensorMapEncodeTiled driver entry point unavailable"); return FM_ERR_NO_DEVICE; }

  IgemmParams p{};
  p.OH = d->OH; p.OW = d->OW; p.B = d->B; p.Bg = Bg;
  const bool resup = d->residual && d->residual_up_h > 0;      // bilinear residual: generic ring only (its box is staged beside it)
  // ---- tile shape: tw*th*tb <= 128 pixels (a tile never straddles two groups)
  int tw = d->tile_w, th = d->tile_h;
  if (tw <= 0 || th <= 0) {
    tw = 1; while (tw < d->OW && tw < 128) tw <<= 1;
    th = 1; while (th < d->OH && tw * th < 128) th <<= 1;
  }
  FM_CHECK_ARG(tw >= 1 && th >= 1 && 128 % (tw * th) == 0, "fm_conv_igemm: tile %dx%d does not divide 128", tw, th);
  int tb = 128 / (tw * th);
  if (G > 1) while (tb > 1 && (tb > Bg || Bg % tb != 0)) tb >>= 1;
  p.tw = tw; p.th = th; p.tb = tb; p.rows = tw * th * tb;
  FM_CHECK_ARG(tw * sx <= 256 && th * sy <= 256 && tb <= 256, "fm_conv_igemm: TMA box too large");
  p.tiles_x = (d->OW + tw - 1) / tw;
  p.tiles_y = (d->OH + th - 1) / th;
  p.tiles_b = (d->B + tb - 1) / tb;
  // ---- split-K for small-M problems: with few (m,n) tiles every CTA would stream the whole K loop of
  // its tile alone; splitting K over idle SMs with the widest N tile divides the per-SM operand bytes
  const int kiters_total = d->ntaps * ((d->Cin + IG_BK - 1) / IG_BK);
  int ksplit = 1;
  {
    static const int env_split = []() { const char* e = getenv("FM3D_SPLITK"); return e ? atoi(e) : 1; }();
    const bool eligible = env_split && nph == 1 && d->tab && !d->colsum && d->residual_up_h <= 0 && d->ksplit != 1 && d->splitk_ws && !d->rgb && !d->border_tab && !d->out_nchw_f32 &&
                          !d->out_cgroup && d->block_n <= 0 && !d->upmode;
    if (eligible) {
      const int bn_wide = d->Cout > 128 ? 256 : (d->Cout > 64 ? 128 : 64);
      const int64_t tiles_wide = static_cast<int64_t>(p.tiles_x) * p.tiles_y * p.tiles_b * ((d->Cout + bn_wide - 1) / bn_wide);
      const int sms = (d->max_ctas > 0 && d->max_ctas < sm_count()) ? d->max_ctas : sm_count();
      int want = d->ksplit > 1 ? d->ksplit : static_cast<int>(sms / (tiles_wide > 0 ? tiles_wide : 1));
      if (want > kiters_total / 6) want = kiters_total / 6;       // keep >= 6 k-iterations per CTA
      if (want > 16) want = 16;
      const int ws_cs = (d->Cout + 15) / 16 * 16;
      const int64_t need = static_cast<int64_t>(d->B) * d->OH * d->OW * ws_cs * 4;
      if (want >= 2 && need <= d->splitk_ws_bytes && (d->ksplit > 1 || tiles_wide * 2 <= sms)) {
        ksplit = want;
        p.ws = d->splitk_ws;
        p.ws_cs = ws_cs;
      }
    }
  }
  // ---- block_n
  const int tbe = d->tab_bstride ? tb : 1;
  int bn = d->block_n;
  if (ksplit > 1) {
    bn = d->Cout > 128 ? 256 : (d->Cout > 64 ? 128 : 64);
  } else if (bn <= 0) {
    bn = d->Cout > 128 ? 256 : (d->Cout > 64 ? 128 : 64);
    // small problems: more, narrower tiles fill more SMs
    const int sms = (d->max_ctas > 0 && d->max_ctas < sm_count()) ? d->max_ctas : sm_count();
    while (bn > 64 && static_cast<int64_t>(p.tiles_x) * p.tiles_y * p.tiles_b * ((d->Cout + bn - 1) / bn) < sms) bn >>= 1;
  }
  if (d->upmode && bn > 128) bn = 128;      // 4 accumulators x BN columns must fit the 512 TMEM columns
  if (ksplit == 1) while (bn > 64 && tbe * bn > IG_TAB_ROWS) bn >>= 1;
  FM_CHECK_ARG(bn == 64 || bn == 128 || bn == 256, "fm_conv_igemm: block_n must be 64/128/256");
  FM_CHECK_ARG(ksplit > 1 || tbe * bn <= IG_TAB_ROWS, "fm_conv_igemm: per-sample tables do not fit (tile_b %d x block_n %d)", tbe, bn);
  FM_CHECK_ARG(!d->out_cgroup || d->out_cgroup % bn == 0 || bn % d->out_cgroup == 0, "fm_conv_igemm: out_cgroup vs block_n");
  p.tiles_n = (d->Cout + bn - 1) / bn;
  p.ksplit = ksplit;
  p.kper = (kiters_total + ksplit - 1) / ksplit;
  p.ksplit = (kiters_total + p.kper - 1) / p.kper;        // no empty k-slices
  const int64_t nt = static_cast<int64_t>(p.tiles_x) * p.tiles_y * p.tiles_b * p.tiles_n * p.ksplit;
  FM_CHECK_ARG(nt < 0x7FFFFFFF, "fm_conv_igemm: too many tiles");
  p.num_tiles = static_cast<int>(nt);
  p.kchunks = (d->Cin + IG_BK - 1) / IG_BK;
  p.ntaps = d->ntaps; p.stride_x = sx; p.stride_y = sy; p.w_rows = d->w_rows; p.Cout = d->Cout;
  p.out = d->out; p.out_H = d->out_H; p.out_W = d->out_W; p.out_cstride = d->out_cstride;
  p.out_y0 = d->out_y0; p.out_x0 = d->out_x0; p.out_ys = d->out_ys; p.out_xs = d->out_xs;
  p.out_nchw_f32 = d->out_nchw_f32;
  FM_CHECK_ARG((d->out_pitch_h == 0 || d->out_pitch_h >= d->out_H) && (d->out_pitch_w == 0 || d->out_pitch_w >= d->out_W) &&
                   (!(d->out_pitch_h || d->out_pitch_w) || !d->out_nchw_f32),
               "fm_conv_igemm: out_pitch_h / out_pitch_w must be >= out_H / out_W (bf16 NHWC outputs only)");
  p.out_PH = d->out_pitch_h > 0 ? d->out_pitch_h : d->out_H;
  p.out_PW = d->out_pitch_w > 0 ? d->out_pitch_w : d->out_W;
  p.tab = d->tab; p.tab_bstride = d->tab_bstride ? 1 : 0;
  p.noise = d->noise; p.noise_bstride = d->noise_bstride ? 1 : 0; p.noise_w = d->noise_w;
  p.residual = static_cast<const __nv_bfloat16*>(d->residual);
  if (d->residual && d->residual_up_h > 0) {
    FM_CHECK_ARG(d->residual_up_w > 0 && d->out_ys == 1 && d->out_xs == 1 && d->out_H > 1 && d->out_W > 1,
                 "fm_conv_igemm: bad upsampled-residual configuration");
    p.res_ih = d->residual_up_h; p.res_iw = d->residual_up_w;
    p.res_ry = static_cast<float>(d->residual_up_h - 1) / (d->out_H - 1);
    p.res_rx = static_cast<float>(d->residual_up_w - 1) / (d->out_W - 1);
  }
  p.rgb = d->rgb;
  p.border_tab = d->border_tab;
  p.out_cgroup = d->out_cgroup; p.out_gstride = d->out_gstride; p.cg_shrink = d->out_cgroup_ow_shrink;
  p.max_ctas = d->max_ctas;
  p.colsum = d->colsum;
  p.tile_ctr = d->tile_counter;       // dropped below when the tile index does not fit the queue's 24 bits
  int max_widx = 0;
  for (int i = 0; i < d->ntaps; ++i) {
    p.tap_dy[i] = d->tap_dy[i]; p.tap_dx[i] = d->tap_dx[i]; p.tap_widx[i] = d->tap_widx[i];
    FM_CHECK_ARG(d->tap_widx[i] >= 0, "fm_conv_igemm: negative tap_widx");
    if (d->tap_widx[i] > max_widx) max_widx = d->tap_widx[i];
  }
  p.nslabs = max_widx + 1;
  p.nbuf = 2;
  if (d->upmode) {
    FM_CHECK_ARG(d->ntaps == 9 && sx == 1 && sy == 1 && G == 1 && d->out_ys == 2 && d->out_xs == 2 && !d->out_nchw_f32 &&
                     !d->out_cgroup && !d->residual && !d->rgb && !d->border_tab && d->OH == d->H + 1 && d->OW == d->W + 1,
                 "fm_conv_igemm: bad upmode configuration");
    p.upmode = 1;
    p.nslabs = 9;
    p.nbuf = (4 * bn * 2 <= 512) ? 2 : 1;
    p.patch_stages = (200 * 1024) / (16384 + 4 * bn * 128);
    if (p.patch_stages > 8) p.patch_stages = 8;
  }
  // ---- row-patch mode: plain 3x3 / stride 1 / pad 1, 128-pixel output rows, N <= 128.  A tile is R
  // output rows (R accumulators in TMEM) so each weight tile loaded from L2 feeds R*128 pixels.
  {
    static const int env_patch = []() { const char* e = getenv("FM3D_PATCH"); return e ? atoi(e) : 1; }();
    static const int env_rows = []() { const char* e = getenv("FM3D_PATCH_ROWS"); return e ? atoi(e) : 0; }();
    bool std33 = d->ntaps == 9 && sx == 1 && sy == 1 && d->x_pixstride == 0 && d->x_rowstride == 0 && d->x_imgstride == 0;
    for (int i = 0; std33 && i < 9; ++i)
      std33 = d->tap_dy[i] == i / 3 - 1 && d->tap_dx[i] == i % 3 - 1 && d->tap_widx[i] == i;
    p.patch = (env_patch && nph == 1 && !resup && std33 && tw == 128 && th == 1 && tb == 1 && bn <= 128 && p.ksplit == 1 && !d->upmode) ? 1 : 0;
    if (p.patch) {
      int R = env_rows > 0 ? env_rows : (bn == 64 ? 4 : 2);
      while (R > 1 && (2 * R * bn > 512 || R > d->OH)) R >>= 1;
      p.prows = R;
      p.patch_a_bytes = (R * 130 * 128 + 1023) & ~1023;
      p.patch_b_bytes = bn * 128;
      p.patch_stage_bytes = p.patch_a_bytes + 3 * bn * 128;
      p.patch_stages = (200 * 1024) / p.patch_stage_bytes;
      if (p.patch_stages > 8) p.patch_stages = 8;
      if (p.patch_stages < 2) { p.patch = 0; p.prows = 1; }
    }
    if (p.patch) {
      p.tiles_y = (d->OH + p.prows - 1) / p.prows;
      p.num_tiles = p.tiles_x * p.tiles_y * p.tiles_b * p.tiles_n;
    } else {
      p.prows = 1;
    }
  }
  // ---- halo-patch mode: plain tap sets (stride 1 or 2) on outputs at least 12 rows tall.  Tile = 8 x 16 output
  // pixels; per channel chunk the input patch the tile's taps touch is loaded ONCE and every tap reads it through a
  // shifted descriptor, so the activation traffic L2 -> SM drops from ntaps x to ~1.4x per chunk.  A strided conv
  // splits its taps by the parity plane of the input they read (iy = s*oy + dy: plane dy mod s, subsampled row
  // oy + (dy - plane)/s); each (chunk, plane) is one patch, fetched with TMA element stride s.
  {
    static const int env_hp = []() { const char* e = getenv("FM3D_HPATCH"); return e ? atoi(e) : 1; }();
    static const int env_hp_s2 = []() { const char* e = getenv("FM3D_HPATCH_S2"); return e ? atoi(e) : 1; }();
    const bool plain_x = d->x_pixstride == 0 && d->x_rowstride == 0 && d->x_imgstride == 0;
    // 0 off | 1 (default) wherever row-patch mode does not apply | 2 always | 3 only N = 256 tiles.  The chip-wide
    // L2 -> SM rate (~6300 B/clk, 42 B/clk per SM) is the budget: a single-CTA tile streams N*128 B of weights
    // per 2N tensor-pipe cycles (64 B/clk) before any activation byte, which is why the row-patch mode shares a
    // weight tile between R accumulators and the CTA-pair mode halves it
    const bool want = env_hp == 2 || (env_hp == 1 && !p.patch) || (env_hp == 3 && !p.patch && bn == 256);
    // stride 2: measured slower than the per-tap scheme for the wide layers (512 -> 3584: 723 -> 959 us; one patch
    // per (chunk, plane) makes 1-tap virtual chunks), faster where the weights stay resident (64 -> 64 at 256^2:
    // 119 -> 97 us) -- so only there (env 2 forces it everywhere)
    const bool s2_resident = bn == 64 && d->Cout <= 64 &&
                             static_cast<int64_t>(kiters_total) * bn * 128 + 3 * 21 * 1024 <= 200 * 1024;
    const bool stride_ok = (sx == 1 && sy == 1) || (sx == 2 && sy == 2 && (env_hp_s2 == 2 || (env_hp_s2 == 1 && s2_resident)));
    // plane / subsampled offset of every tap
    int pl_of[FM_MAX_TAPS], du[FM_MAX_TAPS], dv[FM_MAX_TAPS];
    int du0 = 127, du1 = -127, dv0 = 127, dv1 = -127;
    for (int i = 0; i < d->ntaps; ++i) {
      const int py = ((d->tap_dy[i] % sy) + sy) % sy, px = ((d->tap_dx[i] % sx) + sx) % sx;
      pl_of[i] = py * sx + px;
      du[i] = (d->tap_dy[i] - py) / sy;
      dv[i] = (d->tap_dx[i] - px) / sx;
      du0 = du[i] < du0 ? du[i] : du0; du1 = du[i] > du1 ? du[i] : du1;
      dv0 = dv[i] < dv0 ? dv[i] : dv0; dv1 = dv[i] > dv1 ? dv[i] : dv1;
    }
    // grouped convs: a halo-patch tile lies inside one image, so its group is known per tile (weights are not resident
    // across tiles of different groups: hpw below stays ungrouped)
    static const int env_hp_groups = []() { const char* e = getenv("FM3D_HPATCH_GROUPS"); return e ? atoi(e) : 1; }();
    if (want && !resup && stride_ok && plain_x && (G == 1 || env_hp_groups) && p.ksplit == 1 && !d->upmode && d->ntaps >= 2 && d->OH >= 12 &&
        d->OW >= 8 && dv1 - dv0 <= 8 && du1 - du0 <= 8 && sx * sy <= 4 && (!d->tab_bstride || bn <= IG_TAB_ROWS)) {
      p.hp = 1;
      p.patch = 0; p.prows = 1;
      p.tw = 8; p.th = 16; p.tb = 1; p.rows = 128;
      p.tiles_x = (d->OW + 7) / 8;
      p.tiles_y = (d->OH + 15) / 16;
      p.tiles_b = d->B;
      p.num_tiles = p.tiles_x * p.tiles_y * p.tiles_b * p.tiles_n;
      p.hp_pw = 8 + dv1 - dv0;
      p.hp_ph = 16 + du1 - du0;
      p.hp_dx0 = dv0; p.hp_dy0 = du0;
      p.hp_sx = sx; p.hp_sy = sy;
      p.hp_bytes = (p.hp_pw * p.hp_ph * 128 + 1023) & ~1023;
      // sort the taps by plane (stable) and record each plane's tap range and patch origin
      int order[FM_MAX_TAPS], n = 0, np = 0;
      for (int pl = 0; pl < sx * sy; ++pl) {
        const int first = n;
        for (int i = 0; i < d->ntaps; ++i)
          if (pl_of[i] == pl) order[n++] = i;
        if (n > first) {
          p.hp_pl_first[np] = static_cast<int8_t>(first);
          p.hp_pl_x[np] = static_cast<int16_t>(sx * dv0 + pl % sx);
          p.hp_pl_y[np] = static_cast<int16_t>(sy * du0 + pl / sx);
          ++np;
        }
      }
      p.hp_pl_first[np] = static_cast<int8_t>(n);
      p.hp_np = np;
      for (int k = 0; k < d->ntaps; ++k) {
        const int i = order[k];
        p.tap_dy[k] = d->tap_dy[i]; p.tap_dx[k] = d->tap_dx[i]; p.tap_widx[k] = d->tap_widx[i];
        p.hp_aoff[k] = static_cast<int16_t>(((du[i] - du0) * p.hp_pw + (dv[i] - dv0)) * 8);
      }
      // The patch of virtual chunk c + D is requested when the weights of chunk c start; its slot was last read
      // by chunk c + D - NA (NA = D + E, E >= 2), whose MMAs only need loads issued earlier: no deadlock, and with
      // S <= (E - 1) * T + 1 weight stages (T = taps per virtual chunk) the request does not even block.
      // several phases in one launch: the weight cursor may run S / T chunks ahead of the MMAs, with T the SMALLEST
      // phase (a 1-tap phase: one weight tile per chunk), so more patch slots are kept behind the prefetch distance
      static const int env_up_d = []() { const char* e = getenv("FM3D_UP_D"); return e ? atoi(e) : 2; }();
      static const int env_up_e = []() { const char* e = getenv("FM3D_UP_E"); return e ? atoi(e) : 4; }();
      // FM3D_UP_TCAP: the T of the "weights never run further ahead than the patches" cap below for multi-phase launches.
      // The cap only keeps the patch request from blocking (there is no deadlock for any ring depth); with T = the 1-tap
      // phase the weight ring of the merged up-conv was 4 stages, i.e. ONE chunk of a 4-tap phase in flight.  T = 4: 5
      // stages, 32 -> 64 184 -> 171 us (64 -> 128 unchanged: not ring-depth bound).
      static const int env_up_t = []() { const char* e = getenv("FM3D_UP_TCAP"); return e ? atoi(e) : 4; }();
      const int T = nph > 1 ? (env_up_t > 0 ? env_up_t : ph_tmin) : (d->ntaps + np - 1) / np;
      const int bbytes = bn * 128;
      const int st_max = bn == 256 ? 4 : (bn == 128 ? 6 : 8);
      int D = nph > 1 ? env_up_d : (T >= 5 ? 2 : (T >= 3 ? 3 : 4));
      const int E = nph > 1 ? env_up_e : (T >= 3 ? 2 : 3);
      int st = 0;
      for (; D >= 1; --D) {
        st = (200 * 1024 - (D + E) * p.hp_bytes) / bbytes;
        if (st > st_max) st = st_max;
        if (np == 1 && st > (E - 1) * T + 1) st = (E - 1) * T + 1;
        if (st >= 3 || (D == 1 && st >= 2)) break;
      }
      if (D >= 1 && D + E <= IG_HP_MAXA) {
        p.hp_dist = D; p.hp_na = D + E; p.hp_stages = st; p.hp_bstride = bbytes;
      } else {
        set_error("fm_conv_igemm: halo-patch ring does not fit (patch %d bytes, block_n %d)", p.hp_bytes, bn);
        return FM_ERR_INVALID;
      }
    }
  }
  // ---- clusters: CTAs of a cluster work on adjacent m-tiles of the same n-tile and share each weight
  // tile through one multicast TMA load (weights are the dominant L2->SM stream for N <= 128)
  {
    static const int env_cluster = []() { const char* e = getenv("FM3D_CLUSTER"); return e ? atoi(e) : 2; }();
    p.m_tiles = p.tiles_x * p.tiles_y * p.tiles_b;
    int cs = env_cluster;
    if (cs != 1 && cs != 2 && cs != 4) cs = 1;
    while (cs > 1 && (p.m_tiles % cs != 0 || G > 1 || p.hp)) cs >>= 1;   // no padded m-tiles, one weight slab per cluster
    // CTA pairs (cta_group::2): the two CTAs of a cluster issue ONE M = 256 MMA and each loads only half of the
    // weight tile -- the weight stream is what saturates the L2 -> SM path of a single-CTA tile (64/R B per clk)
    static const int env_pair = []() { const char* e = getenv("FM3D_PAIR"); return e ? atoi(e) : 1; }();
    static const int env_pair_rp = []() { const char* e = getenv("FM3D_PAIR_RP"); return e ? atoi(e) : 1; }();
    p.pair = (env_pair && G == 1 && p.ksplit == 1 && !d->upmode && (!p.patch || env_pair_rp) && p.m_tiles % 2 == 0 &&
              (bn >= 128 || env_pair == 2) && !((d->rgb != nullptr) + (d->residual != nullptr) + (d->border_tab != nullptr) > 1)) ? 1 : 0;
    if (p.pair) cs = 2;
    p.cluster = cs;
    if (p.pair && p.patch) {       // half weight tiles per CTA: smaller stages, deeper ring
      p.patch_b_bytes = bn * 64;
      p.patch_stage_bytes = p.patch_a_bytes + 3 * p.patch_b_bytes;
      p.patch_stages = (200 * 1024) / p.patch_stage_bytes;
      if (p.patch_stages > 8) p.patch_stages = 8;
    }
    // generic ring: a CTA of a pair stages half a weight tile, so its stages are smaller and the ring deeper
    {
      const int sb = IG_BM * IG_BK * 2 + (p.pair ? bn * 64 : bn * 128);
      int st = (200 * 1024) / sb;
      const int st_cfg = bn == 256 ? 4 : (bn == 128 ? 6 : 8);
      if (st > IG_MAX_STAGES) st = IG_MAX_STAGES;
      p.gen_sbytes = p.pair ? sb : IG_BM * IG_BK * 2 + bn * 128;
      p.gen_stages = p.pair ? st : st_cfg;
      if (resup) {
        // two slots of (channel slabs x box pixels) above a shortened operand ring
        FM_CHECK_ARG(G == 1 && p.ksplit == 1 && p.tb == 1, "fm_conv_igemm: an upsampled residual needs an ungrouped, unsplit conv with tiles inside one image");
        const int slab_ch = bn < 128 ? bn : 128;
        p.res_nslab = bn / slab_ch;
        p.res_pitch = (slab_ch + 8) * 2;
        p.res_bw = static_cast<int>(p.res_rx * (p.tw - 1)) + 3;
        p.res_bh = static_cast<int>(p.res_ry * (p.th - 1)) + 3;
        FM_CHECK_ARG(p.res_bw <= 256 && p.res_bh <= 256, "fm_conv_igemm: residual box too large");
        p.res_box_bytes = p.res_pitch * p.res_bw * p.res_bh;
        p.res_slab_bytes = (p.res_box_bytes + 127) & ~127;          // a TMA destination is 128-byte aligned
        p.res_slot_bytes = p.res_nslab * p.res_slab_bytes;
        int stg = (200 * 1024 - 2 * p.res_slot_bytes) / p.gen_sbytes;
        if (stg > p.gen_stages) stg = p.gen_stages;
        FM_CHECK_ARG(stg >= 2, "fm_conv_igemm: residual box (%d bytes) leaves no room for the operand ring", p.res_slot_bytes);
        p.gen_stages = stg;
        p.res_off = (p.gen_stages * p.gen_sbytes + 127) & ~127;
      }
    }
    // halo-patch weight ring, now that the pair decision is known: a CTA of a pair stages only its half of each weight
    // tile, so the same smem holds twice as many stages.  Per-CTA traces of the 64x64 128->128 layer showed the MMA phase
    // of a tile at 2x its tensor-pipe time with ~470 clk per TMA box: the ring depth over the load latency was the limit.
    if (p.hp) {
      static const int env_deep = []() { const char* e = getenv("FM3D_HP_DEEP"); return e ? atoi(e) : 1; }();
      static const int env_up_t2 = []() { const char* e = getenv("FM3D_UP_TCAP"); return e ? atoi(e) : 4; }();
      const int np = p.hp_np, T = nph > 1 ? (env_up_t2 > 0 ? env_up_t2 : ph_tmin) : (d->ntaps + np - 1) / np, E = p.hp_na - p.hp_dist;
      const int bstride = p.pair ? bn * 64 : bn * 128;
      int st = (200 * 1024 - p.hp_na * p.hp_bytes) / bstride;
      if (st > IG_MAX_STAGES) st = IG_MAX_STAGES;
      if (np == 1 && st > (E - 1) * T + 1) st = (E - 1) * T + 1;      // weights never run further ahead than the patches
      if (env_deep && st > p.hp_stages) { p.hp_stages = st; p.hp_bstride = bstride; }
    }
    // weight-resident halo-patch variant: one n-tile and all its (chunk, tap) weight tiles fit beside >= 3 patch slots
    if (p.hp) {
      static const int env_hpw = []() { const char* e = getenv("FM3D_HPW"); return e ? atoi(e) : 1; }();
      const int64_t wbytes = static_cast<int64_t>(kiters_total) * (p.pair ? bn / 2 : bn) * 128;
      int na = static_cast<int>((200 * 1024 - wbytes) / p.hp_bytes);
      if (na > IG_HP_MAXA) na = IG_HP_MAXA;
      static const int env_hpw_min = []() { const char* e = getenv("FM3D_HPW_MIN"); return e ? atoi(e) : 2; }();
      // two patch slots are enough to keep the weights resident (128 -> 128 pair layers: 144 KB of weights + 2 x 23 KB):
      // the next chunk's patch loads while the current chunk's 36 MMAs run
      if (env_hpw && nph == 1 && G == 1 && p.tiles_n == 1 && na >= env_hpw_min) {
        p.hpw = 1;
        p.hp_na = na;
      }
    }
    // weight-resident generic mode (stems with overlapping-window inputs, 1x1 convs): single n-tile, single CTA,
    // all weight tiles + at least 3 A stages fit the ring
    {
      static const int env_wres = []() { const char* e = getenv("FM3D_WRES"); return e ? atoi(e) : 1; }();
      const int64_t wbytes = static_cast<int64_t>(kiters_total) * bn * 128;
      int nst = static_cast<int>((200 * 1024 - wbytes) / (IG_BM * IG_BK * 2));
      const int st_max = bn == 256 ? 4 : (bn == 128 ? 6 : 8);
      if (nst > st_max) nst = st_max;
      if (env_wres && !resup && !p.hp && !p.patch && !p.upmode && !p.pair && p.ksplit == 1 && G == 1 && p.tiles_n == 1 && nst >= 3 &&
          kiters_total >= 2) {
        p.wres = 1;
        p.wres_stages = nst;
        p.cluster = cs = 1;
        // strip sub-mode: 16-byte pixels (x_pixstride = 8 elements: the 3-channel stems), stride 1 in x, taps that differ
        // only in dy, one K chunk (the 8-pixel window).  pSp stem (256^2, 3 taps): 384 TMA rows of 128 bytes per tile
        // became one 6.5 KB box; the layer was TMA-row bound (205 us for 40 us of output writes).
        static const int env_strip = []() { const char* e = getenv("FM3D_STRIP"); return e ? atoi(e) : 1; }();
        bool ok = env_strip && pixs == 8 && rows_ % 8 == 0 && imgs % rows_ == 0 && d->Cin <= 64 && sx == 1 && p.kchunks == 1 && p.tw == 128 &&
                  p.th == 1 && p.tb == 1 && d->ntaps <= 16;
        int dy0 = 127, dy1 = -127;
        for (int i = 0; ok && i < d->ntaps; ++i) {
          ok = d->tap_dx[i] == 0;
          dy0 = d->tap_dy[i] < dy0 ? d->tap_dy[i] : dy0;
          dy1 = d->tap_dy[i] > dy1 ? d->tap_dy[i] : dy1;
        }
        if (ok && dy1 - dy0 + 1 <= 16) {
          p.strip_rows = dy1 - dy0 + 1;
          p.strip_y0 = dy0;
          p.strip_sbytes = (p.strip_rows * 136 * 16 + 127) & ~127;
          int sst = static_cast<int>((200 * 1024 - wbytes) / p.strip_sbytes);
          if (sst > IG_MAX_STAGES) sst = IG_MAX_STAGES;
          if (sst >= 3) {
            p.strip = 1;
            p.strip_k32 = d->Cin <= 32 ? 1 : 0;
            p.wres_stages = sst;
            for (int i = 0; i < d->ntaps; ++i) p.strip_trow[i] = static_cast<int8_t>(d->tap_dy[i] - dy0);
          }
        }
      }
    }
    p.num_super = p.tiles_n * p.ksplit * ((p.m_tiles + cs - 1) / cs);
    p.nph = nph;
    p.nsup1 = p.num_super;
    if (nph > 1) {
      if (!p.hp) {
        set_error("fm_conv_igemm: nphases > 1 needs the halo-patch mode (OH >= 12, OW >= 8, dense NHWC input, FM3D_HPATCH != 0)");
        return FM_ERR_UNSUPPORTED;
      }
      int first = 0;
      for (int i = 0; i < nph; ++i) {
        p.ph_first[i] = static_cast<int8_t>(first);
        p.ph_oy0[i] = static_cast<int16_t>(d->phase_out_y0[i]);
        p.ph_ox0[i] = static_cast<int16_t>(d->phase_out_x0[i]);
        first += d->phase_ntaps[i];
      }
      p.ph_first[nph] = static_cast<int8_t>(first);
      const int64_t tot = static_cast<int64_t>(p.num_super) * nph;
      FM_CHECK_ARG(tot < 0x7FFFFFFF, "fm_conv_igemm: too many tiles");
      p.num_super = static_cast<int>(tot);
    }
  }
  FM_CHECK_ARG(!p.colsum || p.tb == 1, "fm_conv_igemm: colsum needs tiles inside one image (OH*OW >= 128 per image)");
  {
    // whole tiles per cluster only when a cluster gets enough of them for the rounding not to matter (64 -> 128 at B = 32:
    // 7.9 per cluster, 312 -> 292 us; 32 -> 64 with 2.2 per cluster lost 8 %)
    static const int env_whole = []() { const char* e = getenv("FM3D_PHASE_WHOLE"); return e ? atoi(e) : 6; }();
    int sms = sm_count();
    if (d->max_ctas > 0 && d->max_ctas < sms) sms = d->max_ctas;
    const int ncl = sms / p.cluster > 0 ? sms / p.cluster : 1;
    p.ph_whole = (p.nph > 1 && env_whole > 0 && p.nsup1 >= static_cast<int64_t>(env_whole) * ncl && !p.tile_ctr) ? 1 : 0;
  }
  {
    static const int env_lean = []() { const char* e = getenv("FM3D_LEAN_EPI"); return e ? atoi(e) : 1; }();
    p.lean = (env_lean && d->tab && d->out && !d->out_nchw_f32 && !d->out_cgroup && !d->rgb && !d->border_tab && !d->noise && !resup &&
              !p.upmode && !p.patch && p.nph == 1 && p.ksplit == 1 && d->Cout % 16 == 0 && d->out_cstride >= d->Cout &&
              (!d->tab_bstride)) ? 1 : 0;
  }
  if (p.num_super >= (1 << 24) - 1) p.tile_ctr = nullptr;      // queue entries carry 24 bits of tile index
  // ---- tensor maps
  CUtensorMap tmA, tmB;
  {
    const cuuint64_t dims[4] = {static_cast<cuuint64_t>(d->Cin), static_cast<cuuint64_t>(d->W), static_cast<cuuint64_t>(d->H),
                                static_cast<cuuint64_t>(d->B)};
    const cuuint64_t strides[3] = {static_cast<cuuint64_t>(pixs) * 2, static_cast<cuuint64_t>(rows_) * 2,
                                   static_cast<cuuint64_t>(imgs) * 2};
    // with an element stride s TMA loads ceil(box/s) elements: box = n*s loads n
    const cuuint32_t box[4] = {IG_BK, static_cast<cuuint32_t>(p.hp ? p.hp_pw * sx : (p.patch ? 130 : tw * sx)),
                               static_cast<cuuint32_t>(p.hp ? p.hp_ph * sy : (p.patch ? p.prows : th * sy)),
                               static_cast<cuuint32_t>(p.hp ? 1 : tb)};
    const cuuint32_t estr[4] = {1, static_cast<cuuint32_t>(sx), static_cast<cuuint32_t>(sy), 1};
    CUresult r;
    if (p.strip) {
      // the packed image as it is: [B][rows][16-byte pixels][8 channels], dense boxes of strip_rows x 136 pixels
      const cuuint64_t sdims[4] = {8, static_cast<cuuint64_t>(rows_ / 8), static_cast<cuuint64_t>(imgs / rows_), static_cast<cuuint64_t>(d->B)};
      const cuuint32_t sbox[4] = {8, 136, static_cast<cuuint32_t>(p.strip_rows), 1};
      const cuuint32_t one[4] = {1, 1, 1, 1};
      r = encode(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(d->x), sdims, strides, sbox, one,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else
    r = encode(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(d->x), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("fm_conv_igemm: cuTensorMapEncodeTiled(A) failed with CUresult %d", (int)r); return FM_ERR_CUDA; }
  }
  {
    const cuuint64_t dims[2] = {static_cast<cuuint64_t>(d->Cin), static_cast<cuuint64_t>(G) * p.nslabs * d->w_rows};
    const cuuint64_t strides[1] = {static_cast<cuuint64_t>(d->w_cstride) * 2};
    const cuuint32_t box[2] = {IG_BK, static_cast<cuuint32_t>(p.pair ? bn / 2 : bn)};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(d->w), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("fm_conv_igemm: cuTensorMapEncodeTiled(B) failed with CUresult %d", (int)r); return FM_ERR_CUDA; }
  }
  CUtensorMap tmR = tmB;      // unused unless the residual is upsampled
  if (resup) {
    const cuuint64_t dims[4] = {static_cast<cuuint64_t>(d->out_cstride), static_cast<cuuint64_t>(d->residual_up_w),
                                static_cast<cuuint64_t>(d->residual_up_h), static_cast<cuuint64_t>(d->B)};
    const cuuint64_t strides[3] = {static_cast<cuuint64_t>(d->out_cstride) * 2, static_cast<cuuint64_t>(d->out_cstride) * d->residual_up_w * 2,
                                   static_cast<cuuint64_t>(d->out_cstride) * d->residual_up_w * d->residual_up_h * 2};
    const cuuint32_t box[4] = {static_cast<cuuint32_t>(p.res_pitch / 2), static_cast<cuuint32_t>(p.res_bw),
                               static_cast<cuuint32_t>(p.res_bh), 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    FM_CHECK_ARG((reinterpret_cast<uintptr_t>(d->residual) & 15) == 0, "fm_conv_igemm: residual must be 16-byte aligned");
    CUresult r = encode(&tmR, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(d->residual), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("fm_conv_igemm: cuTensorMapEncodeTiled(residual) failed with CUresult %d", (int)r); return FM_ERR_CUDA; }
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  {
    // debug: FM3D_TRACE=1 records per-CTA event clocks of every launch into one device buffer (fm_igemm_trace)
    static const int env_trace = []() { const char* e = getenv("FM3D_TRACE"); return e ? atoi(e) : 0; }();
    if (env_trace) {
      if (!g_trace) cudaMalloc(&g_trace, sizeof(long long) * 1024 * IG_TRACE_N);
      cudaMemsetAsync(g_trace, 0, sizeof(long long) * 1024 * IG_TRACE_N, st);
      p.trace = g_trace;
    }
  }
  switch (bn) {
    case 64: return launch_igemm<64>(tmA, tmB, tmR, p, st);
    case 128: return launch_igemm<128>(tmA, tmB, tmR, p, st);
    default: return launch_igemm<256>(tmA, tmB, tmR, p, st);
  }
}

// Debug: copy the event clocks of the last traced launch (FM3D_TRACE=1) to the host: [n_ctas][slots] int64.
extern "C" int fm_igemm_trace(long long* host_out, int n_ctas, int* slots_out) {
  using namespace fm;
  if (slots_out) *slots_out = IG_TRACE_N;
  FM_CHECK_ARG(host_out && n_ctas > 0 && n_ctas <= 1024, "fm_igemm_trace: bad args");
  if (!g_trace) { set_error("fm_igemm_trace: tracing is off (set FM3D_TRACE=1)"); return FM_ERR_INVALID; }
  FM_CUDA_OK(cudaDeviceSynchronize());
  FM_CUDA_OK(cudaMemcpy(host_out, g_trace, sizeof(long long) * n_ctas * IG_TRACE_N, cudaMemcpyDeviceToHost));
  return FM_OK;
}
