// catan_kernels.cu — sm_100a kernels + C ABI of the vectorised Catan engine (see include/catan_b200.h).
//
// One env step is six launches on three streams, replayed as one CUDA graph (catan_game.cuh holds the game logic; records are
// lane-interleaved chunks of 32 games, and every kernel below works on whole chunks with one game per lane):
//
//   caller's stream
//   transition_kernel    one block per chunk, the HOT range of the chunk staged in shared memory (TMA bulk copy).  The games are
//                        sorted by action type and the four warps take eight each: translate + validate + the scalar part of
//                        apply_action, ONE THREAD PER GAME.  Then warp 0: the incremental longest-road update of the games that
//                        placed a road / settlement (t_lr_fast: ~93 % are settled by two tiny walks); the other warps: the
//                        data-parallel follow-ups (dice payout, belief updates), one warp per game and one lane per item.  A
//                        game whose longest road needs a real search, or that ended, is queued and copied into a staging chunk.
//   encode_kernel<ROLE_ROWS>   one block of four warps per chunk: the packed observation rows, built as bit images in registers
//                        and written as 256-bit streaming stores.  Queued games are left out.
//   encode_kernel<ROLE_MASKS>  one block of four warps per chunk: done / reward / info, the legal-action masks as bit sets (the
//                        board scans behind the placement masks one warp per game that needs one) and the fused random-legal sampler.
//
//   library's high-priority stream 1 (forked after the transition, joined at the end of the step)
//   lr_slow_kernel       one 512-thread block per queued update: the paths through the new road -- or, when the stored
//                        length cannot be trusted, the reference's full enumeration (game.py:843-862) -- as a pool of
//                        16-byte subtree tasks that the lanes drain and re-split without barriers (lp_pool).
//   encode_kernel<LISTED> the whole encode (eight warps) for the searched games, on their staging chunks; each block copies its
//                        games home at the end, and the last block of the launch banks and clears the queue counters.
//
//   library's high-priority stream 2
//   encode_kernel<LISTED> the same for the games that ended in this step: done / reward, Board.reset + Game.reset (serial per
//                        game, hence off the main path), rows and masks of the new game; one game per block.
//
// Why the search is not inside a thread-per-game kernel: its cost varies by four orders of magnitude between games and
// would stall 31 other games per unit of imbalance; why it runs on its own stream: it is latency-bound (a few hundred
// dependent walk steps) and touches 0.3 % of the games.  (Since round 2 it no longer hides behind the encode of the others:
// profiles/r2_notes.md.)
#include <cuda_runtime.h>

#include "device_scope.cuh"

#include <cstddef>
#include <cstdlib>
#include <new>
#include <string>
#include <vector>

#include "../../include/catan_b200.h"
#include "catan_game.cuh"

namespace catanb {

__device__ const Topo d_topo = CATAN_TOPO_INITIALIZER;

// -DCATAN_PROFILE_PHASES (profiles/phase_profile.py only): cycles from block start to a few markers, summed over the blocks
#ifdef CATAN_PROFILE_PHASES
// (d_phase is defined in catan_game.cuh, where the reset is instrumented as well)
#define CATAN_MARK(k_) do { if ((threadIdx.x & 31) == 0) { atomicAdd(&d_phase[2 * (k_)], static_cast<unsigned long long>(clock64() - t_block0)); atomicAdd(&d_phase[2 * (k_) + 1], 1ull); } } while (0)
#define CATAN_MARK_BEGIN() const long long t_block0 = clock64()
#else
#define CATAN_MARK(k_) ((void)0)
#define CATAN_MARK_BEGIN() ((void)0)
#endif

// ---- launch shapes ------------------------------------------------------------------------------
constexpr int kTransWarps = 4;              // transition_kernel: warps per chunk of 32 games ...
#ifndef CATAN_LR_BATCH
#define CATAN_LR_BATCH 6
#endif
constexpr int kLrBatch = CATAN_LR_BATCH;                 // incremental longest road: games walked at a time per rule warp
constexpr int kTransThreads = kTransWarps * 32;
constexpr int kRowWarps = CATAN_OBS_TILE_PARTS;  // encode_kernel: warps that write observation rows (a tile part and a player part each) ...
constexpr int kEncWarps = 2 * kRowWarps;         // ... and as many that share the per-game scalar work (done / reward, masks, sampler)
constexpr int kEncThreads = kEncWarps * 32;
constexpr int kMaskWarps = kEncWarps - kRowWarps;
// the rows launch of a step (ROLE_ROWS): its own number of warps, each with a balanced share of the row's 60 32-byte pairs of pieces
#ifndef CATAN_ROWS_WARPS
#define CATAN_ROWS_WARPS 4
#endif
#ifndef CATAN_ROWS_BLOCKS
#define CATAN_ROWS_BLOCKS 7
#endif
constexpr int kRowsWarps = CATAN_ROWS_WARPS, kRowsThreads = kRowsWarps * 32;
constexpr int kMasksThreads = kMaskWarps * 32;
static_assert(kRowsWarps >= 4 && kRowsWarps <= 8, "warps 0-3 of the rows launch own the four player blocks");
// Warp w of the rows launch writes the player part 4 + w (w < 4: five 32-byte pairs of pieces) or the card lists (w == min(4, kRowsWarps - 1):
// four pairs) and the pairs [rows_tile_lo(w), rows_tile_lo(w + 1)) of the 36 pairs of the header + tile region, so that every warp
// has about 60 / kRowsWarps pairs.
__host__ __device__ constexpr int rows_fixed_pairs(int w) { return (w < 4 ? 5 : 0) + (w == (kRowsWarps > 4 ? 4 : 3) ? 4 : 0); }
__host__ __device__ constexpr int rows_tile_lo(int w) {
  // greedy: hand out the 36 tile pairs one by one to the warp with the fewest pairs so far (ties: the highest warp)
  int have[12] = {0};
  for (int k = 0; k < kRowsWarps; ++k) have[k] = rows_fixed_pairs(k);
  int tiles[12] = {0};
  for (int n = 0; n < 36; ++n) {
    int best = kRowsWarps - 1;
    for (int k = kRowsWarps - 1; k >= 0; --k) if (have[k] < have[best]) best = k;
    have[best] += 1; tiles[best] += 1;
  }
  int lo = 0;
  for (int k = 0; k < w; ++k) lo += tiles[k];
  return lo;
}
static_assert(rows_tile_lo(0) == 0 && rows_tile_lo(kRowsWarps) == 36, "the tile pairs are covered exactly once");
static_assert(CATAN_OBS_PARTS == 2 * CATAN_OBS_TILE_PARTS + 1, "row warp r writes tile part r and player part r; the last one also the lists");
constexpr int kLrSlowThreads = 512;         // lr_slow_kernel: one block per update that needs a search
constexpr int kLrSlowBlocksPerSM = 2;
constexpr int kSampleThreads = 128;         // stand-alone sampler kernel

enum { MODE_STEP = 0, MODE_RESET = 1, MODE_REFRESH = 2 };

struct LrCtl {                // per-step queue counters; queue_finish() banks them into the totals and clears them
  int32_t count, slow_count;  // longest-road updates of this step; those that went to lr_slow_kernel (= length of its queue)
  int32_t rs_count, pad_;     // games that ended in this step and are reset on their own stream (= length of rs_queue)
  int32_t blocks_done[2];     // blocks of the two queues' encode launches that are through: the last one banks and clears its queue's counters
  int32_t search_claim, pad2_;   // next unclaimed entry of the search queue (lr_slow_kernel's blocks take the searches one by one)
  unsigned long long total, slow_total, rs_total;   // the same, summed over all earlier steps
  unsigned long long dbg[6];  // lr_slow_kernel diagnostics: cycles sum / max, walk steps sum / max per search; full enumerations; tasks
  unsigned long long hist[3][24];   // log2 histograms: cycles of a search, walk steps of a search, cycles of the LONGEST search of a step
  unsigned long long step_max;      // (longest search of the current step; banked into hist[2] by queue_finish)
};

struct EnvParams {
  uint8_t* recs;               // lane-interleaved chunks of 32 games (catan_game.cuh)
  uint8_t* stage;              // same layout: compact copy of the games with a pending longest-road update, in queue order
  int n_envs;
  uint64_t seed, first_env_id;
  catan_config_t cfg;
  int list_kind;               // LISTED encode: 0 = the searched games (lr queue), 1 = the games that ended (reset queue)
  const int32_t* actions;      // transition input
  int32_t* actions_out;        // fused sampler output (may alias `actions`), or nullptr
  uint8_t* obs;
  uint8_t* masks;
  float* reward;
  uint8_t* info;
  uint32_t* err_flags;
  const uint8_t* env_mask;     // envs whose byte is 0 are left untouched; nullptr = all envs
  int range_first, range_count;   // env range this launch covers
  uint32_t* side;              // [n] transition -> encode: err | acted_pid << 8 | act_type << 16 | roll << 24 | longest-road update pending << 31
  uint64_t* lr_slow_queue;     // [n] updates that need a search: env index | PlayerId << 32 | edge or corner << 40 | CATAN_LR_* << 48 | acting PlayerId << 56
  LrCtl* lr_ctl;               // queue lengths of this step
  uint64_t* rs_queue;          // [n] env index of the games that ended in this step (auto-reset): they are finished, reset and encoded on
  uint8_t* stage_rs;           //     staging copies by their own launch, so that the slow, serial reset never holds up a block of 32 games
  // what a LISTED encode / copy-back launch works on: one of the two queues
  const uint64_t* list_queue;
  uint8_t* list_stage;
  const int32_t* list_count;
  int list_group;              // games of the queue per block (a power of two <= 32)
};

struct GameSmem {
  Topo topo;
  TopoX topox;
};

__device__ __forceinline__ void stage_topology(GameSmem& S, int tid, int nthreads) {
  const int4* src = reinterpret_cast<const int4*>(&d_topo);
  int4* dst = reinterpret_cast<int4*>(&S.topo);
  for (int i = tid; i < static_cast<int>(sizeof(Topo) / 16); i += nthreads) dst[i] = src[i];
  build_topox(d_topo, S.topox, tid, nthreads);
  __syncthreads();
}

// The chunk of 32 games a block works on is staged in shared memory: a game's fields are spread over the whole 26 KB
// chunk, and the rule code reads them through long chains of dependent loads -- from HBM/L2 one miss (0.3-1 us) at a
// time, a block needed ~100 us for a few thousand instructions.  The copy is one burst of independent 16-byte loads.
// Both directions go through the TMA engine as ONE 1-D bulk copy issued by one thread (cp.async.bulk, SASS UBLKCP):
// with ordinary 16-byte loads / stores the two copies were 36 % of the transition kernel's stall samples.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar) {           // by one thread; make it visible with a block barrier
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(bar)) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// by one thread: start the copy of a chunk into shared memory; everybody then waits with chunk_wait(bar, phase)
__device__ __forceinline__ void bulk_to_shared(uint8_t* dst, const uint8_t* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void chunk_to_shared(uint8_t* dst, const uint8_t* chunk, uint64_t* bar) {
  bulk_to_shared(dst, chunk, static_cast<uint32_t>(CATAN_CHUNK_BYTES), bar);
}
__device__ __forceinline__ void chunk_wait(uint64_t* bar, uint32_t phase) {
  uint32_t ok;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(phase) : "memory");
  } while (!ok);
}
// every thread that wrote to the staged chunk: chunk_written(), then a block barrier; then ONE thread: chunk_to_global()
__device__ __forceinline__ void chunk_written() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_to_global(uint8_t* dst, const uint8_t* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the block may exit (and its shared memory go) after this
}
__device__ __forceinline__ void chunk_to_global(uint8_t* chunk, const uint8_t* src) { bulk_to_global(chunk, src, static_cast<uint32_t>(CATAN_CHUNK_BYTES)); }

// The games that need a search are scattered over the chunks: every field access of a warp that works on 32 of them
// would touch 32 sectors.  The transition therefore copies them (one warp per game) into staging chunks in queue order;
// they are searched and encoded there with the ordinary coalesced code and copied back.
__device__ __forceinline__ void copy_game(const GameView& src, const GameView& dst, int lane) {
  constexpr int n16 = static_cast<int>(offsetof(GameRec, rng_ctr) / 2), b0 = static_cast<int>(offsetof(GameRec, robber_tile)), nb = static_cast<int>(sizeof(GameRec)) - b0;
  int16_t h[(n16 + 31) / 32];
  uint8_t c[(nb + 31) / 32];
#pragma unroll
  for (int q = 0; q < (n16 + 31) / 32; ++q) if (lane + 32 * q < n16) h[q] = src.raw<int16_t>(2 * (lane + 32 * q));    // all loads first
#pragma unroll
  for (int q = 0; q < (nb + 31) / 32; ++q) if (lane + 32 * q < nb) c[q] = src.raw<uint8_t>(b0 + lane + 32 * q);
  const uint32_t w = lane < 3 ? src.raw<uint32_t>(static_cast<int>(offsetof(GameRec, rng_ctr)) + 4 * lane) : 0u;
  const uint16_t v = lane < 2 ? src.raw<uint16_t>(static_cast<int>(offsetof(GameRec, actions_this_turn)) + 2 * lane) : 0;
#pragma unroll
  for (int q = 0; q < (n16 + 31) / 32; ++q) if (lane + 32 * q < n16) dst.raw<int16_t>(2 * (lane + 32 * q)) = h[q];
#pragma unroll
  for (int q = 0; q < (nb + 31) / 32; ++q) if (lane + 32 * q < nb) dst.raw<uint8_t>(b0 + lane + 32 * q) = c[q];
  if (lane < 3) dst.raw<uint32_t>(static_cast<int>(offsetof(GameRec, rng_ctr)) + 4 * lane) = w;
  if (lane < 2) dst.raw<uint16_t>(static_cast<int>(offsetof(GameRec, actions_this_turn)) + 2 * lane) = v;
}

// ---- 1. transition ------------------------------------------------------------------------------
// One block per chunk of 32 games.  Warp 0 runs the scalar part of apply_action, one game per lane; the data-parallel
// follow-ups it posts (dice payout over 19 tiles x 6 corners, belief updates over 60 entries) are then executed by all
// warps of the block, one warp per game and one lane per item.
constexpr uint32_t kHotBytes = (CATAN_HOT_END - CATAN_HOT_BEGIN) * 32;     // 5 920 of the 26 624 bytes of a chunk
template <bool DIRECT>
struct alignas(128) TransSmem {
  uint8_t chunk[DIRECT ? ((kHotBytes + 127) / 128) * 128 : CATAN_CHUNK_BYTES];   // DIRECT: only the hot range of the chunk is staged
  uint64_t mbar;
  alignas(16) Topo topo;
  StepTmp tmp[32];
  int32_t n_follow;
  int32_t slot[32];          // staging slot of a game that goes to lr_slow_kernel, else -1
  uint8_t order[kTransWarps][32];   // the chunk's games sorted by action type (one copy per warp: no barrier needed)
  uint64_t und[kLrBatch][54];   // incremental longest road: neighbour tables of the games being walked
  uint8_t follow_list[32];
  int follow_claim;           // next unclaimed entry of follow_list
};

// DIRECT = false stages the whole chunk in shared memory (TMA in, TMA out).  DIRECT = true stages only its HOT range (GameRec: flags,
// hands, bank, points, counters -- 185 of the 832 bytes of a record, what the rules of every action read and write) and touches the
// rest (board, beliefs, card lists) in place through L1 / L2: the rules use a small, data-dependent part of those, and with 11 KB
// instead of 32 KB of shared memory per block twice as many chunks are in flight per SM.
#ifndef CATAN_TRANS_DIRECT_BLOCKS
#define CATAN_TRANS_DIRECT_BLOCKS 16
#endif
template <bool DIRECT>
__global__ void __launch_bounds__(kTransThreads, DIRECT ? CATAN_TRANS_DIRECT_BLOCKS : 7) transition_kernel(const __grid_constant__ EnvParams P) {
  __shared__ TransSmem<DIRECT> S;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  CATAN_MARK_BEGIN();
  const int base = (P.range_first & ~31) + static_cast<int>(blockIdx.x) * 32;
  uint8_t* const home = P.recs + static_cast<size_t>(base >> 5) * CATAN_CHUNK_BYTES;
  // view of game b of this chunk: hot fields in the staged copy, the others in the staged copy too (whole chunk) or at home (DIRECT)
  uint8_t* const hot = DIRECT ? S.chunk : S.chunk + static_cast<size_t>(CATAN_HOT_BEGIN) * 32;
  uint8_t* const coldp = DIRECT ? home : S.chunk;
#define CATAN_TVIEW(b_) GameView(hot, (b_), coldp)
  if (tid == 0) {                                                    // in flight while the topology is staged
    mbar_init(&S.mbar);
    if (DIRECT) bulk_to_shared(S.chunk, home + static_cast<size_t>(CATAN_HOT_BEGIN) * 32, kHotBytes, &S.mbar);
    else chunk_to_shared(S.chunk, home, &S.mbar);
    S.n_follow = 0; S.follow_claim = 0;
  }
  {
    const int4* src = reinterpret_cast<const int4*>(&d_topo);
    int4* dst = reinterpret_cast<int4*>(&S.topo);
    for (int k = tid; k < static_cast<int>(sizeof(Topo) / 16); k += kTransThreads) dst[k] = src[k];
  }
  // order of the chunk's games by action type (frozen / out-of-range games last), computed by every warp for itself
  const int ig = base + lane;
  const bool live = ig >= P.range_first && ig < P.range_first + P.range_count && !(P.env_mask != nullptr && P.env_mask[ig] == 0);
  int n_live;
  {
    int key = 14;
    if (live) { const int ty = P.actions[static_cast<size_t>(ig) * CATAN_ACTION_WORDS + CATAN_A_TYPE]; key = (ty < 0 || ty > 12) ? 13 : ty; }
    int pos = 0, cnt = 0;
#pragma unroll
    for (int t = 0; t < 15; ++t) {
      const unsigned b = __ballot_sync(0xffffffffu, key == t);
      if (key == t) pos = cnt + __popc(b & ((1u << lane) - 1u));
      if (t == 13) n_live = cnt + __popc(b);
      cnt += __popc(b);
    }
    S.order[warp][pos] = static_cast<uint8_t>(lane);
    if (warp == 0) {
      S.slot[lane] = -1;
      if (!live) { S.tmp[lane].lr_pid = 0; S.tmp[lane].err = 0; S.tmp[lane].follow = 0; }
    }
  }
  __syncthreads();
  chunk_wait(&S.mbar, 0);
  if (warp == 0) CATAN_MARK(0);
  // The rule code is one big switch over 13 action types, and a warp pays for every type that occurs among ITS games: with one
  // game per lane of one warp the 32 games of a chunk took 22 us of a 38 us block (profiles/r2_notes.md).  The games were
  // therefore sorted by action type above (S.order), and each of the four warps takes eight consecutive games of that order:
  // mostly one or two types per warp, and the four warps run side by side.
  const int slot8 = warp * 8 + lane;
  const bool mine = lane < 8 && slot8 < n_live;
  const int gl = mine ? S.order[warp][slot8] : 0;                    // game of this thread inside the chunk
  const int i = base + gl;
  TCx cx;
  cx.g = CATAN_TVIEW(gl);
  cx.T = &S.topo; cx.X = nullptr; cx.cfg = &P.cfg; cx.seed = P.seed; cx.env_id = P.first_env_id + static_cast<uint64_t>(i);   // (X: masks only)
  {
    bool follow = false;
    if (mine) {
      cx.s = load_seats(cx.g);
      StepTmp& tmp = S.tmp[gl];
      t_step_scalar(cx, P.actions + static_cast<size_t>(i) * CATAN_ACTION_WORDS, tmp);
      if (tmp.err) P.err_flags[i] |= 1u << tmp.err;
      follow = tmp.follow != 0;
    }
    const unsigned fb = __ballot_sync(0xffffffffu, follow);
    int pos = 0;
    if (lane == 0 && fb) pos = atomicAdd(&S.n_follow, __popc(fb));
    pos = __shfl_sync(0xffffffffu, pos, 0);
    if (follow) S.follow_list[pos + __popc(fb & ((1u << lane) - 1u))] = static_cast<uint8_t>(gl);
    CATAN_MARK(1);
  }
  __syncthreads();
  if (warp == 0) {
    // (game = lane from here on)
    cx.g = CATAN_TVIEW(lane);
    const int i = base + lane;
    const bool lr = live && !S.tmp[lane].err && S.tmp[lane].lr_pid;
    // longest road (game.py:843-919), one thread per update: the incremental rule settles ~93 % of them on the spot; a
    // game that needs a search goes to the queue of lr_slow_kernel.  (Independent of the follow-ups: those touch hands,
    // bank and beliefs only.)
    const StepTmp& tmp = S.tmp[lane];
    const unsigned lb = __ballot_sync(0xffffffffu, lr);
    // Two games at a time: the warp gathers the road / blocked-corner bit sets (one lane per corner / edge) and the
    // neighbour table of the player's road graph (one lane per corner) into shared memory, then the two owning lanes walk.
    bool slow = false;
    for (unsigned mm = lb; mm;) {
      int mine_k = -1;
      RoadBits rb = {0ull, 0ull, 0u};
#pragma unroll
      for (int k = 0; k < kLrBatch; ++k) {
        if (!mm) break;
        const int b = __ffs(static_cast<int>(mm)) - 1;
        mm &= mm - 1;
        const uint32_t pid = __shfl_sync(0xffffffffu, static_cast<uint32_t>(tmp.lr_pid), b);
        const GameView gb = CATAN_TVIEW(b);
        const uint32_t c0 = gb.corner(lane), c1 = lane + 32 < 54 ? gb.corner(lane + 32) : 0u;
        const uint32_t k0 = __ballot_sync(0xffffffffu, c0 != 0 && (c0 >> 2) != pid), k1 = __ballot_sync(0xffffffffu, c1 != 0 && (c1 >> 2) != pid);
        const uint32_t e0 = __ballot_sync(0xffffffffu, gb.edge(lane) == pid), e1 = __ballot_sync(0xffffffffu, gb.edge(lane + 32) == pid);
        const uint32_t e2 = __ballot_sync(0xffffffffu, lane + 64 < 72 && gb.edge(lane + 64 < 72 ? lane + 64 : 0) == pid);
        const RoadBits r = {e0 | (static_cast<uint64_t>(e1) << 32), k0 | (static_cast<uint64_t>(k1) << 32), e2};
        S.und[k][lane] = t_road_nb(S.topo, r, lane);
        if (lane + 32 < 54) S.und[k][lane + 32] = t_road_nb(S.topo, r, lane + 32);
        if (lane == b) { rb = r; mine_k = k; }
      }
      __syncwarp();
      if (mine_k >= 0) {
        const int len = t_lr_fast(cx.g, S.topo, tmp.lr_pid, tmp.lr_kind, tmp.lr_loc, tmp.acted_pid, &rb, S.und[mine_k]);
        if (len >= 0) t_lr_apply(cx.g, tmp.lr_pid, len, false, nullptr);
        else slow = true;
      }
      __syncwarp();
    }
    if (slow) {
      const int slot = atomicAdd(&P.lr_ctl->slow_count, 1);
      P.lr_slow_queue[slot] = static_cast<uint64_t>(static_cast<uint32_t>(i)) | (static_cast<uint64_t>(tmp.lr_pid) << 32) | (static_cast<uint64_t>(tmp.lr_loc) << 40) |
                              (static_cast<uint64_t>(tmp.lr_kind) << 48) | (static_cast<uint64_t>(tmp.acted_pid) << 56);
      S.slot[lane] = slot;
    }
    if (lane == 0 && lb) atomicAdd(&P.lr_ctl->count, __popc(lb));
    // A game that ends with this step is reset by the encode (wrapper.py:30-34 after done): a handful of serial shuffles with a
    // rejection loop, 60-300 us for ONE lane while the other 319 threads of an encode block wait -- the ~40 such blocks per step
    // were the stragglers that set the encode kernel's duration (profiles/r2_notes.md).  The points are final here (follow-ups do
    // not touch them), so these games leave the main path like the searched ones and go to the reset queue.
    bool ends = false;
    if (live && !slow && !tmp.err && P.cfg.auto_reset)
      ends = cx.g.vp(0) >= 10 || cx.g.vp(1) >= 10 || cx.g.vp(2) >= 10 || cx.g.vp(3) >= 10;
    if (ends) {
      const int slot = atomicAdd(&P.lr_ctl->rs_count, 1);
      P.rs_queue[slot] = static_cast<uint64_t>(static_cast<uint32_t>(i));
      S.slot[lane] = -2 - slot;
    }
    if (live)
      P.side[i] = static_cast<uint32_t>(tmp.err) | (static_cast<uint32_t>(tmp.acted_pid) << 8) | (static_cast<uint32_t>(tmp.act_type) << 16) |
                  (static_cast<uint32_t>(tmp.roll_info) << 24) | ((slow || ends) ? 0x80000000u : 0u);
    CATAN_MARK(2);
  }
  {
    // the follow-ups of the chunk's games, one warp per game: every warp claims the next one (warp 0 joins when its longest-road
    // updates are done; the slowest warp of the slowest block is what the kernel waits for)
    const int nf = S.n_follow;
    for (;;) {
      int j = 0;
      if (lane == 0) j = atomicAdd(&S.follow_claim, 1);
      j = __shfl_sync(0xffffffffu, j, 0);
      if (j >= nf) break;
      const GameView g = CATAN_TVIEW(S.follow_list[j]);
      t_followups_group(g, S.topo, S.tmp[g.lane], lane, 32);
    }
    if (warp != 0) CATAN_MARK(3);
  }
  chunk_written();
  __syncthreads();
  for (int b = warp; b < 32; b += kTransWarps) {                     // games that wait for a search: also into their staging slot
    const int slot = S.slot[b];
    if (slot >= 0) copy_game(CATAN_TVIEW(b), game_view(P.stage, static_cast<size_t>(slot)), lane);
    else if (slot <= -2) copy_game(CATAN_TVIEW(b), game_view(P.stage_rs, static_cast<size_t>(-2 - slot)), lane);
  }
  if (tid == 0) {                                                    // (frozen games go back unchanged)
    if (DIRECT) bulk_to_global(home + static_cast<size_t>(CATAN_HOT_BEGIN) * 32, S.chunk, kHotBytes);
    else chunk_to_global(home, S.chunk);
  }
#undef CATAN_TVIEW
  if (warp == 0) CATAN_MARK(4);
}

// ---- 2. longest road ----------------------------------------------------------------------------
// Measured on random play at ticks 1000-1600 (oracle, 38 k road placements): 3.5 % of the steps trigger an update; the
// reference's full enumeration visits 205 corners on average and 29 k at most.  96 % of the new roads are bridges of the
// player's road graph, where the incremental rule of t_lr_fast() (inside transition_kernel) is exact with two walks of
// ~10 visits; 166 updates per step (65 536 games) come here: one BLOCK per update, see lp_pool in catan_core.cuh.
constexpr int kLrRing = 2048;               // task ring of the pool (shared memory; a power of two >= 4 x lanes + 64)
struct alignas(16) LrSmem {
  Topo topo;
  uint64_t adj[54], adjb[54];
  LpTask ring[kLrRing];
  uint8_t paths[54 * kLrSlowThreads];
  int32_t best[4];
  int32_t ctl[CATAN_LP_CTL_WORDS];
  int32_t steps, tasks;      // diagnostics: walk steps and tasks of the current update
  int32_t job;               // the queue entry this block works on
};

__device__ __forceinline__ int block_longest_path(LrSmem& S, const GameView& g, int pid, int tid) {
  if (tid < 32) t_lp_build_adj(g, S.topo, pid, S.adj, tid, 32);
  if (tid == 0) S.best[0] = 0;
  __syncthreads();
  CATAN_LP_RUN(S.adj, S.adj, -1, -1, S.ctl, S.best, S.paths, kLrSlowThreads, tid, S.ring, kLrRing, tid == 0, __syncthreads());
  if (tid == 0) { S.steps += S.ctl[3]; S.tasks += S.ctl[1]; }
  const int r = S.best[0];
  __syncthreads();
  return r;
}
// longest path that contains the new road `edge` (block-cooperative twin of t_through_edge)
__device__ __forceinline__ int block_through_edge(LrSmem& S, const GameView& g, int pid, int edge, int tid) {
  int through = 0;
  for (int dir = 0; dir < 2; ++dir) {
    const int a = S.topo.edge_corners[edge][dir], b = S.topo.edge_corners[edge][dir ^ 1];
    const uint8_t bd = g.corner(a);
    if (bd && (bd >> 2) != pid) continue;                            // (block-uniform) a is blocked: no arc a -> b
    if (tid < 32) t_lp_build_adj2(g, S.topo, pid, S.adj, S.adjb, b, tid, 32);
    if (tid == 0) S.best[0] = 0;
    __syncthreads();
    CATAN_LP_RUN(S.adj, S.adjb, b, a, S.ctl, S.best, S.paths, kLrSlowThreads, tid, S.ring, kLrRing, tid == 0, __syncthreads());
    if (tid == 0) { S.steps += S.ctl[3]; S.tasks += S.ctl[1]; }
    through = max(through, S.best[0]);
    __syncthreads();
  }
  return through;
}

#ifndef CATAN_LR_MIN_BLOCKS
#define CATAN_LR_MIN_BLOCKS 1
#endif
__global__ void __launch_bounds__(kLrSlowThreads, CATAN_LR_MIN_BLOCKS) lr_slow_kernel(const __grid_constant__ EnvParams P) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  LrSmem& S = *reinterpret_cast<LrSmem*>(smem_raw);
  const int tid = threadIdx.x;
  const int count = P.lr_ctl->slow_count;
  if (static_cast<int>(blockIdx.x) >= count) return;
  {
    const int4* src = reinterpret_cast<const int4*>(&d_topo);
    int4* dst = reinterpret_cast<int4*>(&S.topo);
    for (int i = tid; i < static_cast<int>(sizeof(Topo) / 16); i += kLrSlowThreads) dst[i] = src[i];
  }
  __syncthreads();
  // A steady-state step queues about as many searches as there are resident blocks (~290 of 296) and their lengths differ by an
  // order of magnitude: the blocks CLAIM the searches one by one instead of striding over the queue.
  for (;;) {
    __syncthreads();
    if (tid == 0) S.job = atomicAdd(&P.lr_ctl->search_claim, 1);
    __syncthreads();
    const int j = S.job;
    if (j >= count) break;
    const uint64_t en = P.lr_slow_queue[j];
    const GameView g = game_view(P.stage, static_cast<size_t>(j));
    const int pid = static_cast<int>((en >> 32) & 0xff), loc = static_cast<int>((en >> 40) & 0xff), kind = static_cast<int>((en >> 48) & 0xff);
    const long long t_start = clock64();
    bool was_full = true;
    if (tid == 0) { S.steps = 0; S.tasks = 0; }
    // The search reads a few dozen fields of the game one after the other (adjacency build, stored lengths, t_lr_apply), each a
    // round trip to L2 on the staging chunk: 8-33 us of a typical search were this latency.  All threads touch the record's bytes
    // once, side by side, so that those reads hit L1.
    {
      unsigned acc = 0;
      for (int k = tid; k < static_cast<int>(sizeof(GameRec)); k += kLrSlowThreads) acc += g.raw<uint8_t>(k & ~0) ;
      if (acc == 0xffffffffu) S.tasks = -1;                          // (keeps the loads)
    }
    __syncthreads();
    if (kind == CATAN_LR_ROAD && loc != 0xff && !g.lr_dirty(pid - 1)) {
      was_full = false;
      // the stored length is exact: only the paths through the new road can beat it
      const int old = g.lr_holder() == pid ? g.lr_count() : (g.has_path_key(pid - 1) ? g.cur_longest_path(pid - 1) : 0);
      const int through = block_through_edge(S, g, pid, loc, tid);
      if (tid == 0) t_lr_apply(g, pid, max(through, old), false, nullptr);
    } else {
      const int len = block_longest_path(S, g, pid, tid);
      const bool shrunk = t_lr_is_shrunk(g, pid, len);               // game.py:880-881: re-measure the other three players
      uint8_t other[5] = {0, 0, 0, 0, 0};
      if (shrunk)
        for (int o = WHITE; o <= RED; ++o)
          if (o != pid) other[o] = static_cast<uint8_t>(block_longest_path(S, g, o, tid));
      if (tid == 0) { t_lr_apply(g, pid, len, shrunk, other); g.lr_dirty(pid - 1) = 0; }
    }
    __syncthreads();                                                 // everybody has read the game before its next update
    if (tid == 0) {
      const unsigned long long dt = static_cast<unsigned long long>(clock64() - t_start);
      atomicAdd(&P.lr_ctl->dbg[0], dt); atomicMax(&P.lr_ctl->dbg[1], dt);
      atomicAdd(&P.lr_ctl->dbg[2], static_cast<unsigned long long>(S.steps)); atomicMax(&P.lr_ctl->dbg[3], static_cast<unsigned long long>(S.steps));
      if (was_full) atomicAdd(&P.lr_ctl->dbg[4], 1ull);
      atomicAdd(&P.lr_ctl->hist[0][min(23, 63 - __clzll(static_cast<long long>(dt | 1ull)))], 1ull);
      atomicAdd(&P.lr_ctl->hist[1][min(23, 31 - __clz(S.steps | 1))], 1ull);
      atomicMax(&P.lr_ctl->step_max, dt);
      atomicAdd(&P.lr_ctl->dbg[5], static_cast<unsigned long long>(S.tasks));
    }
  }
}

// by the LAST block of a queue's encode launch (every consumer of the queue counters has run): bank and clear them
__device__ __forceinline__ void queue_finish(LrCtl* c, int kind) {
  if (kind == 0) {
    c->total += static_cast<unsigned long long>(c->count); c->slow_total += static_cast<unsigned long long>(c->slow_count);
    c->count = 0; c->slow_count = 0; c->search_claim = 0;
    if (c->step_max) { c->hist[2][min(23, 63 - __clzll(static_cast<long long>(c->step_max)))] += 1ull; c->step_max = 0ull; }
  } else {
    c->rs_total += static_cast<unsigned long long>(c->rs_count);
    c->rs_count = 0;
  }
}
__device__ __forceinline__ void queue_block_done(LrCtl* c, int kind, int tid) {   // all threads of a block of a LISTED encode launch, at its end
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    if (atomicAdd(&c->blocks_done[kind], 1) == static_cast<int>(gridDim.x) - 1) {
      __threadfence();
      queue_finish(c, kind);
      c->blocks_done[kind] = 0;
    }
  }
}

// ---- 3. finish + masks + sampler + observation --------------------------------------------------
// One block per chunk of 32 games, one game per lane in every warp.  Warp 0: done / reward / info (+ auto-reset), then
// the legal-action masks and the fused sampler.  Warps 1..CATAN_OBS_PARTS: one 16-byte aligned piece of the observation
// row each (t_obs_part_lo): the header + tile pieces as bit sets expanded in registers, the player blocks and card
// lists through a 128-byte window per thread.  A block's latency is what bounds the kernel (a warp issues one instruction
// every ~40 cycles), hence the many narrow parts.
struct alignas(128) EncSmem {
  uint8_t chunk[CATAN_CHUNK_BYTES];                                 // the 32 games of this block (see chunk_to_shared)
  uint64_t mbar;
  GameSmem topo;
  uint32_t scan_need, reset_need;
  int32_t scan_pid_done;                                            // (profiling build: warps that finished)
  // ---- a ROLE_ROWS block needs nothing below
  Scan scan[32];                                                    // board scan of game b (valid where scan_need has bit b)
  uint8_t scan_pid[32];
  // ---- a ROLE_MASKS block of a step (no resets on the main path) needs nothing below
  uint32_t wbuf[CATAN_RESET_WORDS];                                 // reset: pre-drawn Philox words (warp 0)
  alignas(4) uint8_t arr[128];                                       // reset: scratch (warp 0)
};
// A step's main encode is TWO launches of 4-warp blocks, because both halves are instruction-fetch bound (no_inst was the first or
// second stall reason of the active warps, profiles/r2_notes.md: the code of a block runs once per chunk, with little reuse) and
// neither waits for the other: ROLE_ROWS writes the observation rows, ROLE_MASKS does done / reward / info, masks and sampler.
// Each then has half the code, half the registers per block and 7 blocks per SM.  Every other launch (reset, refresh, the two
// queues) is ROLE_BOTH: 8 warps with block-wide barriers.
enum { ROLE_BOTH = 0, ROLE_MASKS = 1, ROLE_ROWS = 2 };
constexpr size_t enc_smem_bytes(int role) {
  return role == ROLE_ROWS ? offsetof(EncSmem, scan) : role == ROLE_MASKS ? offsetof(EncSmem, wbuf) : sizeof(EncSmem);
}

// LISTED = false: block b takes chunk b of the env range; in a step, games whose longest-road update is still pending
// (side bit 31) are left out.  LISTED = true: those games, 32 per block iteration in queue order, on their staging
// copies once the searches have finished (second stream, see launch_step).
#ifndef CATAN_ENC_MIN_BLOCKS
#define CATAN_ENC_MIN_BLOCKS 4   // (5 blocks per SM at 48 registers measured 2 % slower than 4 at 56)
#endif
template <int MODE, bool SAMPLE, bool LISTED, int ROLE>
__global__ void __launch_bounds__(ROLE == ROLE_BOTH ? kEncThreads : ROLE == ROLE_ROWS ? kRowsThreads : kMasksThreads,
                                  ROLE == ROLE_BOTH ? CATAN_ENC_MIN_BLOCKS : ROLE == ROLE_ROWS ? CATAN_ROWS_BLOCKS : 7)
encode_kernel(const __grid_constant__ EnvParams P) {
  static_assert(ROLE == ROLE_BOTH || (MODE == MODE_STEP && !LISTED), "the split roles are for the main path of a step");
  constexpr int kThreads = ROLE == ROLE_BOTH ? kEncThreads : ROLE == ROLE_ROWS ? kRowsThreads : kMasksThreads, kWarps = kThreads / 32;
  extern __shared__ __align__(128) uint8_t enc_smem_raw[];
  EncSmem& S = *reinterpret_cast<EncSmem*>(enc_smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  CATAN_MARK_BEGIN();
  const int list_count = LISTED ? *P.list_count : 0;
  const int G = LISTED ? P.list_group : 32;                          // games per block iteration: LISTED blocks take a WINDOW of a staging chunk
  if (LISTED && static_cast<int>(blockIdx.x) * G >= list_count) { queue_block_done(P.lr_ctl, P.list_kind, tid); return; }
#define CATAN_ENC_HOME(l0_) ((LISTED ? P.list_stage : P.recs) + static_cast<size_t>((LISTED ? (l0_) : (P.range_first & ~31) + (l0_)) >> 5) * CATAN_CHUNK_BYTES)
  if (tid == 0) {                                                    // the first chunk is on its way while the topology is staged
    S.reset_need = 0; S.scan_need = 0; S.scan_pid_done = 0;          // (made visible by the barrier of stage_topology)
    mbar_init(&S.mbar);
    chunk_to_shared(S.chunk, CATAN_ENC_HOME((static_cast<int>(blockIdx.x) * G) & ~31), &S.mbar);
  }
  stage_topology(S.topo, tid, kThreads);
  uint32_t phase = 0;
  for (int w0 = static_cast<int>(blockIdx.x) * G; LISTED ? w0 < list_count : w0 == static_cast<int>(blockIdx.x) * G; w0 += static_cast<int>(gridDim.x) * G) {
    const int l0 = w0 & ~31;                                         // first queue slot / game of the chunk
    int i;
    bool valid;
    if (LISTED) {
      valid = l0 + lane >= w0 && l0 + lane < w0 + G && l0 + lane < list_count;
      i = valid ? static_cast<int>(static_cast<uint32_t>(P.list_queue[l0 + lane])) : 0;
    } else {
      i = (P.range_first & ~31) + l0 + lane;
      valid = i >= P.range_first && i < P.range_first + P.range_count && !(MODE != MODE_REFRESH && P.env_mask != nullptr && P.env_mask[i] == 0);
      if (MODE == MODE_STEP && valid) valid = !(P.side[i] >> 31);
    }
    // the chunk of these 32 games (home records, or the staging copies) -> shared memory; lane b of every warp works on
    // game b of it.  What the block changes goes home explicitly: the games of the other stream must not be touched.
    uint8_t* const home = CATAN_ENC_HOME(l0);
    if (tid == 0 && w0 != static_cast<int>(blockIdx.x) * G) chunk_to_shared(S.chunk, home, &S.mbar);
    chunk_wait(&S.mbar, phase);
    phase ^= 1;
    if (!LISTED && warp == 0) CATAN_MARK(8);
#define CATAN_VIEW_OF(b_) GameView{S.chunk, (b_)}
    // Roles (kMaskWarps + kRowWarps warps).
    //   mask warps 0 .. kMaskWarps-1: the per-game scalar work -- done / reward / info, the legal-action masks, the sampler.  It is a
    //     long dependent chain per game, and when ONE warp did it for all 32 games it was the block's long pole (34 us of a ~40 us
    //     block, profiles/r2_notes.md); mask warp q takes the games g with g % kMaskWarps == q in its lanes g / kMaskWarps.
    //   row warps: lane b works on game b; row warp r writes tile part r of the observation row and then player part r (the last
    //     one also the card lists).
    // In a step launch no game of the main path is ever reset (the transition sent those to the reset queue), and nothing the
    // observation reads is changed by done / reward: the row warps then start at once and the mask warps synchronise among
    // themselves (named barrier 1).  Every other launch (reset, refresh, the two queues) keeps the block-wide barriers.
    constexpr bool kDecoupled = MODE == MODE_STEP && !LISTED;
    const bool mask_warp = ROLE == ROLE_MASKS || (ROLE == ROLE_BOTH && warp < kMaskWarps);
    const int mq = warp;                                             // index among the mask warps
    const int gm = lane * kMaskWarps + mq;                           // game of this thread in its mask role
    const bool m_lane = mask_warp && gm < 32;
    const int im = __shfl_sync(0xffffffffu, i, gm & 31);
    const bool vm = __shfl_sync(0xffffffffu, static_cast<int>(valid), gm & 31) != 0 && m_lane;
#define CATAN_MASK_SYNC() do { if (kDecoupled && ROLE == ROLE_BOTH) asm volatile("bar.sync 1, %0;" :: "n"(kMaskWarps * 32) : "memory"); else __syncthreads(); } while (0)
    if (ROLE != ROLE_ROWS && (mask_warp || !kDecoupled)) {
      TCx mx;                                                        // context of the mask role
      mx.g = CATAN_VIEW_OF(gm & 31);
      mx.T = &S.topo.topo; mx.X = &S.topo.topox; mx.cfg = &P.cfg; mx.seed = P.seed; mx.env_id = P.first_env_id + static_cast<uint64_t>(im);
      const GameView hv = GameView{home, gm & 31};
      uint8_t* info = P.info + static_cast<size_t>(im) * CATAN_INFO_STRIDE;
      if (mask_warp && MODE != MODE_REFRESH) {
        bool need_reset = false;
        if (MODE == MODE_STEP) {
          if (vm) {
            mx.s = load_seats(mx.g);
            const uint32_t sd = P.side[im];
            StepTmp tmp;
            tmp.err = static_cast<uint8_t>(sd); tmp.acted_pid = static_cast<uint8_t>(sd >> 8); tmp.act_type = static_cast<uint8_t>(sd >> 16);
            tmp.roll_info = static_cast<uint8_t>((sd >> 24) & 0x7f);
            need_reset = (ROLE == ROLE_MASKS || LISTED) ? t_step_finish_inl(mx, tmp, P.reward + static_cast<size_t>(im) * 4, info)
                                              : t_step_finish(mx, tmp, P.reward + static_cast<size_t>(im) * 4, info);
            hv.episode_steps() = mx.g.episode_steps(); hv.winner() = mx.g.winner();   // what done / reward changed (wrapper.py:85-112)
#pragma unroll
            for (int p = 0; p < 4; ++p) hv.curr_vps(p) = mx.g.curr_vps(p);
          }
        } else {
          need_reset = vm;
        }
        if (need_reset && !kDecoupled) atomicOr(&S.reset_need, 1u << gm);
      }
      if (!LISTED && warp == 0) CATAN_MARK(9);
      if (!kDecoupled) {
        __syncthreads();
        if (MODE != MODE_REFRESH && S.reset_need) {                  // (block-uniform)
          // Board.reset + Game.reset: warp 0 does them game by game (the lanes share the shuffles' draws and the 6 / 8 check)
          if (warp == 0) {
            unsigned rb = S.reset_need;
#ifdef CATAN_PROFILE_PHASES
            const long long t_reset0 = clock64();
#endif
            while (rb) {
              const int b = __ffs(static_cast<int>(rb)) - 1;
              rb &= rb - 1;
              const int e = __shfl_sync(0xffffffffu, i, b);
              reset_game_group(CATAN_VIEW_OF(b), S.topo.topo, P.seed, P.first_env_id + static_cast<uint64_t>(e), S.wbuf, S.arr,
                               lane, 32, MODE == MODE_STEP ? P.info + static_cast<size_t>(e) * CATAN_INFO_STRIDE : nullptr);
              copy_game(CATAN_VIEW_OF(b), GameView{home, b}, lane);  // the whole new game goes home
            }
#ifdef CATAN_PROFILE_PHASES
            if (lane == 0 && MODE == MODE_STEP) {
              const unsigned long long d = static_cast<unsigned long long>(clock64() - t_reset0);
              atomicAdd(&d_phase[32], d); atomicAdd(&d_phase[33], 1ull); atomicMax(&d_phase[48], d);
              atomicAdd(&d_phase[50 + (d < 20000 ? 0 : d < 40000 ? 1 : d < 80000 ? 2 : d < 160000 ? 3 : d < 320000 ? 4 : 5)], 1ull);
            }
#endif
          }
          __syncthreads();
        }
      }
      // (the games are final)
      if (mask_warp) {
        MaskBits m;
        MaskPlan pl;
        pl.post = 0;
        if (vm) mx.s = load_seats(mx.g);
        if (MODE != MODE_STEP && vm) t_write_info_fresh(mx.g, info, MODE == MODE_RESET);
        const bool need_scan = vm && ((ROLE == ROLE_MASKS || LISTED) ? t_masks_pre_inl(mx, m, pl) : t_masks_pre(mx, m, pl));
        if (m_lane) S.scan_pid[gm] = static_cast<uint8_t>(mx.g.players_go());
        if (need_scan) atomicOr(&S.scan_need, 1u << gm);
        if (!LISTED && warp == 0) CATAN_MARK(10);
        CATAN_MASK_SYNC();
        // the board scans (54 corners, 72 edges, 19 tiles) of the games that are in a placement phase: one WARP per game, one
        // lane per corner / edge; the mask warps (in a step) or all warps take their share
        {
          unsigned nb = S.scan_need;
          for (int k = 0; nb; ++k) {
            const int b = __ffs(static_cast<int>(nb)) - 1;
            nb &= nb - 1;
            if (k % (kDecoupled ? kMaskWarps : kWarps) != warp) continue;
            const Scan r = t_scan_group(CATAN_VIEW_OF(b), S.topo.topo, S.topo.topox, S.scan_pid[b], lane, 32);
            if (lane == 0) S.scan[b] = r;
          }
        }
        CATAN_MASK_SYNC();
        if (!LISTED && warp == 0) CATAN_MARK(11);
        if (vm) {
          if (pl.post) { if (ROLE == ROLE_MASKS || LISTED) t_masks_post_inl(mx, m, pl, S.scan[gm]); else t_masks_post(mx, m, pl, S.scan[gm]); }
          {
            MaskFlat F;
            t_flatten_masks(m, F);
            t_store_mask_row(F, P.masks + static_cast<size_t>(im) * CATAN_MASK_STRIDE);
          }
          if (SAMPLE) {
            const int ap = t_current_actor(mx.g) - 1;
            uint32_t hand = 0;
#pragma unroll
            for (int r = 0; r < 5; ++r) hand |= static_cast<uint32_t>(mx.g.res(ap, r) != 0) << r;
            const uint32_t decision = mx.g.decision_ctr();
            hv.decision_ctr() = decision + 1;
            if (ROLE == ROLE_MASKS || LISTED) t_sample_action_inl(m, hand, P.seed, mx.env_id, decision, P.actions_out + static_cast<size_t>(im) * CATAN_ACTION_WORDS);
            else t_sample_action(m, hand, P.seed, mx.env_id, decision, P.actions_out + static_cast<size_t>(im) * CATAN_ACTION_WORDS);
          }
        }
        if (!LISTED) CATAN_MARK(12);
      } else {                                                       // row warps of a launch with block-wide barriers: their share of the scans
        __syncthreads();
        {
          unsigned nb = S.scan_need;
          for (int k = 0; nb; ++k) {
            const int b = __ffs(static_cast<int>(nb)) - 1;
            nb &= nb - 1;
            if (k % kWarps != warp) continue;
            const Scan r = t_scan_group(CATAN_VIEW_OF(b), S.topo.topo, S.topo.topox, S.scan_pid[b], lane, 32);
            if (lane == 0) S.scan[b] = r;
          }
        }
        __syncthreads();
      }
    }
    if (ROLE != ROLE_MASKS && !mask_warp) {
      TCx cx;
      cx.g = CATAN_VIEW_OF(lane);
      cx.T = &S.topo.topo; cx.X = &S.topo.topox; cx.cfg = &P.cfg; cx.seed = P.seed; cx.env_id = P.first_env_id + static_cast<uint64_t>(i);
      if (valid) {
        cx.s = load_seats(cx.g);
        uint8_t* row = P.obs + static_cast<size_t>(i) * CATAN_OBS_STRIDE;
        if constexpr (ROLE == ROLE_ROWS) {                           // a balanced share of the row's pieces (rows_tile_lo)
          int lo = 0, hi = 0;
#pragma unroll
          for (int w = 0; w < kRowsWarps; ++w) if (w == warp) { lo = 32 * rows_tile_lo(w); hi = 32 * rows_tile_lo(w + 1); }
          if (hi > lo) t_encode_obs_tiles_inl<true>(cx, row, lo, hi);
          CATAN_MARK(13);
          if (warp < 4) t_encode_obs_part<true>(cx, row, CATAN_OBS_TILE_PARTS + warp);
          if (warp == (kRowsWarps > 4 ? 4 : 3)) t_encode_obs_part<true>(cx, row, CATAN_OBS_PARTS - 1);
        } else {
          const int r = warp - kMaskWarps;                           // tile part r, then player part r (the last row warp: the lists too)
          if (LISTED) t_encode_obs_tiles_inl<false>(cx, row, t_obs_part_lo(r), t_obs_part_lo(r + 1));   // (few blocks, latency matters: inlined)
          else t_encode_obs_part(cx, row, r);
          if (!LISTED) CATAN_MARK(13);
          t_encode_obs_part(cx, row, CATAN_OBS_TILE_PARTS + r);
          if (r == kRowWarps - 1) t_encode_obs_part(cx, row, CATAN_OBS_PARTS - 1);
        }
      }
      if (!LISTED) CATAN_MARK(14);
    }
#undef CATAN_MASK_SYNC
#ifdef CATAN_PROFILE_PHASES
    if (!LISTED) {                                                   // the last warp of the block to get here: the block's duration
      __syncwarp();
      int last = 0;
      if (lane == 0) last = atomicAdd(&S.scan_pid_done, 1) == kWarps - 1;
      if (last) { CATAN_MARK(15); const unsigned long long d = static_cast<unsigned long long>(clock64() - t_block0); atomicMax(&d_phase[40], d);
                  atomicAdd(&d_phase[42 + (d < 30000 ? 0 : d < 60000 ? 1 : d < 90000 ? 2 : d < 120000 ? 3 : d < 200000 ? 4 : 5)], 1ull);
                  if (S.reset_need) atomicAdd(&d_phase[41], 1ull); }
    }
#endif
    if (LISTED) {
      __syncthreads();                                               // the block's changes to the staging copies are all written
      for (int b = warp; b < 32; b += kWarps) {                      // staging -> home records, one warp per game
        const int vb = __shfl_sync(0xffffffffu, static_cast<int>(valid), b), ib = __shfl_sync(0xffffffffu, i, b);
        if (vb) copy_game(GameView{home, b}, game_view(P.recs, static_cast<size_t>(ib)), lane);
      }
      if (tid == 0) { S.reset_need = 0; S.scan_need = 0; }           // the scratch and the flags are reused by the next 32 games
      __syncthreads();
    }
#undef CATAN_VIEW_OF
  }
#undef CATAN_ENC_HOME
  if (LISTED) queue_block_done(P.lr_ctl, P.list_kind, tid);
}

// Game.randomise_uncertainty (game.py:1207-1282) for every env whose byte in `controlling` is a PlayerId: one thread per game on
// the home records (a few hundred draws and byte moves per game; this is the forward-search hook, not the step path).  A game whose
// beliefs admit no consistent deal within max_attempts gets bit CATAN_ERR_NO_DEAL in its sticky error word.
__global__ void __launch_bounds__(128) randomise_kernel(const __grid_constant__ EnvParams P, const uint8_t* controlling, int max_attempts) {
  const int e = blockIdx.x * 128 + threadIdx.x;
  if (e >= P.n_envs) return;
  const int c = controlling[e];
  if (c < WHITE || c > RED) return;
  TCx cx;
  cx.g = game_view(P.recs, static_cast<size_t>(e));
  cx.T = &d_topo; cx.X = nullptr; cx.cfg = &P.cfg; cx.seed = P.seed; cx.env_id = P.first_env_id + static_cast<uint64_t>(e);
  cx.s = load_seats(cx.g);
  if (t_randomise_uncertainty(cx, c, max_attempts) == 0) P.err_flags[e] |= 1u << CATAN_ERR_NO_DEAL;
}

// the compact host transport of action rows (catan_step_sample_host_async_u8): one byte per word, 255 = -1
__global__ void __launch_bounds__(256) actions_unpack_kernel(const uint8_t* __restrict__ in, int32_t* __restrict__ out, int n_words) {
  const int k = (blockIdx.x * 256 + threadIdx.x) * 4;
  if (k >= n_words) return;                                          // (n_words is a multiple of 4: 20 words per env)
  const uchar4 b = *reinterpret_cast<const uchar4*>(in + k);
  *reinterpret_cast<int4*>(out + k) = make_int4(b.x == 255 ? -1 : b.x, b.y == 255 ? -1 : b.y, b.z == 255 ? -1 : b.z, b.w == 255 ? -1 : b.w);
}
__global__ void __launch_bounds__(256) actions_pack_kernel(const int32_t* __restrict__ in, uint8_t* __restrict__ out, int n_words) {
  const int k = (blockIdx.x * 256 + threadIdx.x) * 4;
  if (k >= n_words) return;
  const int4 v = *reinterpret_cast<const int4*>(in + k);
  auto pk = [](int x) { return static_cast<unsigned char>(x < 0 ? 255 : (x > 254 ? 254 : x)); };
  *reinterpret_cast<uchar4*>(out + k) = make_uchar4(pk(v.x), pk(v.y), pk(v.z), pk(v.w));
}

// stand-alone sampler: one thread per env, reads the bound mask / obs rows back from global memory
__global__ void __launch_bounds__(kSampleThreads) sample_kernel(uint8_t* recs, int n_envs, uint64_t seed, uint64_t first_env_id,
                                                                const uint8_t* masks, const uint8_t* obs, int32_t* actions_out) {
  const int e = blockIdx.x * kSampleThreads + threadIdx.x;
  if (e >= n_envs) return;
  const GameView g = game_view(recs, static_cast<size_t>(e));
  MaskBits m;
  t_load_mask_row(masks + static_cast<size_t>(e) * CATAN_MASK_STRIDE, m);
  const uint8_t* cur = obs + static_cast<size_t>(e) * CATAN_OBS_STRIDE + CATAN_OBS_CURRENT_RES + 1;
  uint32_t hand = 0;
#pragma unroll
  for (int r = 0; r < 5; ++r) hand |= static_cast<uint32_t>(cur[r] != 0) << r;
  const uint32_t dec = g.decision_ctr();
  g.decision_ctr() = dec + 1;
  t_sample_action(m, hand, seed, first_env_id + static_cast<uint64_t>(e), dec, actions_out + static_cast<size_t>(e) * CATAN_ACTION_WORDS);
}

}  // namespace catanb

// =================================================================================================
// host side: handle + C ABI
// =================================================================================================
using catanb::EnvParams;
using catanb::GameRec;

static thread_local std::string g_last_error;
static int fail(const std::string& msg) { g_last_error = msg; return -1; }
extern "C" void catan_set_last_error(const char* msg) { g_last_error = msg ? msg : ""; }   // used by ppo_kernels.cu
static int cuda_fail(cudaError_t e, const char* what) { return fail(std::string(what) + ": " + cudaGetErrorString(e)); }
#define CATAN_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return cuda_fail(e_, #call); } while (0)

struct catan_env {
  int n = 0, device = 0, sm_count = 0;
  uint64_t seed = 0, first_env_id = 0;
  catan_config_t cfg{};
  uint8_t* recs = nullptr;            // ceil(n / 32) lane-interleaved chunks
  uint8_t* stage = nullptr;           // staging chunks of the games with a pending longest-road update
  size_t rec_bytes = 0;
  size_t window_bytes = 0;            // L2 access-policy window on the records (0: not available)
  uint32_t* err_flags = nullptr;
  uint32_t* side = nullptr;
  uint64_t* lr_slow_queue = nullptr;
  uint64_t* rs_queue = nullptr;       // games that ended in the step (auto-reset): env indices
  uint8_t* stage_rs = nullptr;        // their staging chunks
  catanb::LrCtl* lr_ctl = nullptr;
  int32_t* actions_stage = nullptr;   // device staging for catan_step_host
  uint8_t* actions_u8_stage = nullptr;   // device staging of the one-byte-per-word host transport (allocated on first use)
  uint8_t* obs = nullptr;
  uint8_t* masks = nullptr;
  float* reward = nullptr;
  uint8_t* info = nullptr;
  int lr_grid = 0;
  cudaStream_t lr_stream = nullptr;   // high-priority stream of the longest-road updates (overlaps the encode kernel)
  cudaStream_t rs_stream = nullptr;   // high-priority stream of the games that are reset (likewise)
  cudaStream_t rows_stream = nullptr; // the observation rows of a step: beside the masks + sampler launch (neither reads what the other writes)
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_join_rs = nullptr, ev_join_rows = nullptr;
  // catan_set_timing: CUDA events around the two kernels on the caller's stream, a ring of kTimedSteps steps
  // catan_set_graphs: every distinct step call (entry point + buffer pointers) is captured once into a CUDA graph on an internal
  // stream and replayed on the caller's stream afterwards: one driver call per step instead of ~20 (9 launches, 6 event calls, copies)
  bool trans_direct = true;           // transition_kernel<DIRECT> (see there)
  size_t enc_pad_bytes = 0;           // extra dynamic shared memory of the rows / masks launches: caps their blocks per SM (room for the search blocks)
  bool rows_beside = false;           // the rows launch of a step on the library's rows stream, beside the masks launch
  bool use_graphs = false;
  cudaStream_t capture_stream = nullptr;
  struct StepGraph { int kind; const void* p[4]; cudaGraphExec_t exec; };
  std::vector<StepGraph> graphs;
  bool timing = false;
  cudaEvent_t tev[32][8] = {};        // [0] transition [1] rows [3] masks + sampler [2]; search stream: [1] wait [4] search [5] rest [6]; reset stream: .. [7]
  unsigned long long timed = 0;        // steps recorded since timing was switched on
  double t_ms[8] = {};                 // transition, encode (rows + masks), rows alone, [3] fork -> search starts, [4] search, [5] the rest of
                                       // the search stream, [6] fork -> end of the reset stream, [7] fork -> the step's last join: summed over the steps already retired from the ring
  unsigned long long t_n = 0;
};

// the handle's device for the rest of the calling function; the caller's current device comes back when it returns
#define CATAN_ON_DEVICE_OF(env_)                                              \
  catanb::DeviceScope device_scope_;                                          \
  if (device_scope_.enter((env_)->device)) return fail("cannot make the handle's device current")

static EnvParams make_params(const catan_env* env) {
  EnvParams P{};
  P.recs = env->recs; P.stage = env->stage; P.n_envs = env->n; P.seed = env->seed; P.first_env_id = env->first_env_id; P.cfg = env->cfg;
  P.obs = env->obs; P.masks = env->masks; P.reward = env->reward; P.info = env->info; P.err_flags = env->err_flags;
  P.side = env->side; P.lr_slow_queue = env->lr_slow_queue; P.lr_ctl = env->lr_ctl;
  P.rs_queue = env->rs_queue; P.stage_rs = env->stage_rs;
  return P;
}

static int game_blocks(int first, int count) {   // chunks of 32 games touched by [first, first + count): one block each
  return ((first + count - 1) >> 5) - (first >> 5) + 1;
}

// With CATAN_L2_WINDOW set in the environment the two big kernels are launched with an L2 access-policy window on the game
// records (persisting on hit, streaming on miss): the 54 MB of records then survive in the 126 MB L2 from the transition to
// the encode and to the next step, while the 150 MB of observation / mask rows written per step stream through
// (st.global.cs).  Measured: -1 % step time; off by default because the persisting set-aside is taken from every other
// user of the device's L2 (the policy network).
template <class Kernel>
static cudaError_t launch_with_record_window(const catan_env* env, Kernel kernel, int blocks, int threads, size_t smem, cudaStream_t stream,
                                             const EnvParams& P) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(blocks)); cfg.blockDim = dim3(static_cast<unsigned>(threads));
  cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeAccessPolicyWindow;
  attr[0].val.accessPolicyWindow.base_ptr = env->recs;
  attr[0].val.accessPolicyWindow.num_bytes = env->window_bytes;
  attr[0].val.accessPolicyWindow.hitRatio = 1.0f;
  attr[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
  attr[0].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
  cfg.attrs = attr; cfg.numAttrs = env->window_bytes ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, P);
}

template <int MODE, bool SAMPLE>
static int launch_encode(catan_env* env, EnvParams P, int first, int count, cudaStream_t stream, cudaEvent_t* tev = nullptr) {
  P.range_first = first; P.range_count = count;
  if (count <= 0) return 0;
  if constexpr (MODE == catanb::MODE_STEP) {
    // Two launches of 4-warp blocks, side by side: the observation rows on the library's rows stream (forked at ev_fork, joined by
    // the caller), masks + sampler on the caller's stream.  With the timing hooks on they run one after the other on the caller's
    // stream, an event between them, so that each is measured alone.
    cudaStream_t rs = (tev || !env->rows_beside) ? stream : env->rows_stream;
    if (rs != stream) CATAN_CUDA(cudaStreamWaitEvent(rs, env->ev_fork, 0));
    CATAN_CUDA(launch_with_record_window(env, catanb::encode_kernel<catanb::MODE_STEP, SAMPLE, false, catanb::ROLE_ROWS>, game_blocks(first, count),
                                         catanb::kRowsThreads, catanb::enc_smem_bytes(catanb::ROLE_ROWS) + env->enc_pad_bytes, rs, P));
    if (tev) CATAN_CUDA(cudaEventRecord(tev[3], stream));
    if (rs != stream) CATAN_CUDA(cudaEventRecord(env->ev_join_rows, rs));
    CATAN_CUDA(launch_with_record_window(env, catanb::encode_kernel<catanb::MODE_STEP, SAMPLE, false, catanb::ROLE_MASKS>, game_blocks(first, count),
                                         catanb::kMasksThreads, catanb::enc_smem_bytes(catanb::ROLE_MASKS) + env->enc_pad_bytes, stream, P));
    if (rs != stream) CATAN_CUDA(cudaStreamWaitEvent(stream, env->ev_join_rows, 0));
  } else {
    CATAN_CUDA(launch_with_record_window(env, catanb::encode_kernel<MODE, false, false, catanb::ROLE_BOTH>,
                                         game_blocks(first, count), catanb::kEncThreads, sizeof(catanb::EncSmem), stream, P));
  }
  return 0;
}

// One step.  On the caller's stream: transition, then the encode of every game whose longest road is settled (96 %).
// On the library's high-priority stream, forked after the transition and joined at the end: the queued longest-road
// updates and the encode of exactly those games.  The searches are latency-bound (a few hundred dependent walk steps
// per update) and would otherwise sit between the two big kernels with the machine idle.
static int retire_timed_step(catan_env* env, cudaEvent_t* ev) {     // one ring slot -> the sums (waits for that step)
  float a = 0.f, b = 0.f, c = 0.f;
  CATAN_CUDA(cudaEventSynchronize(ev[2]));
  CATAN_CUDA(cudaEventElapsedTime(&a, ev[0], ev[1]));
  CATAN_CUDA(cudaEventElapsedTime(&b, ev[1], ev[2]));
  CATAN_CUDA(cudaEventElapsedTime(&c, ev[1], ev[3]));
  env->t_ms[0] += a; env->t_ms[1] += b; env->t_ms[2] += c; env->t_n += 1;
  CATAN_CUDA(cudaEventSynchronize(ev[6]));
  CATAN_CUDA(cudaEventSynchronize(ev[7]));
  float d = 0.f;
  CATAN_CUDA(cudaEventElapsedTime(&d, ev[1], ev[4])); env->t_ms[3] += d;
  CATAN_CUDA(cudaEventElapsedTime(&d, ev[4], ev[5])); env->t_ms[4] += d;
  CATAN_CUDA(cudaEventElapsedTime(&d, ev[5], ev[6])); env->t_ms[5] += d;
  CATAN_CUDA(cudaEventElapsedTime(&d, ev[1], ev[7])); env->t_ms[6] += d;
  { float e6 = 0.f, e7 = 0.f; CATAN_CUDA(cudaEventElapsedTime(&e6, ev[1], ev[6])); CATAN_CUDA(cudaEventElapsedTime(&e7, ev[1], ev[7]));
    env->t_ms[7] += fmaxf(fmaxf(e6, e7), b); }
  return 0;
}

template <bool SAMPLE>
static int launch_step(catan_env* env, EnvParams P, cudaStream_t stream) {
  P.range_first = 0; P.range_count = env->n;
  cudaEvent_t* tev = nullptr;
  if (env->timing) {
    tev = env->tev[env->timed % 32];
    if (env->timed >= 32 && retire_timed_step(env, tev)) return -1;
    env->timed += 1;
    CATAN_CUDA(cudaEventRecord(tev[0], stream));
  }
  if (env->trans_direct) CATAN_CUDA(launch_with_record_window(env, catanb::transition_kernel<true>, game_blocks(0, env->n), catanb::kTransThreads, 0, stream, P));
  else CATAN_CUDA(launch_with_record_window(env, catanb::transition_kernel<false>, game_blocks(0, env->n), catanb::kTransThreads, 0, stream, P));
  if (tev) CATAN_CUDA(cudaEventRecord(tev[1], stream));
  CATAN_CUDA(cudaEventRecord(env->ev_fork, stream));
  CATAN_CUDA(cudaStreamWaitEvent(env->lr_stream, env->ev_fork, 0));
  CATAN_CUDA(cudaStreamWaitEvent(env->rs_stream, env->ev_fork, 0));
  {   // the searched games: search -> encode on the staging copies (8 games of the queue per block) -> home
    EnvParams L = P;
    L.list_queue = env->lr_slow_queue; L.list_stage = env->stage; L.list_count = &env->lr_ctl->slow_count; L.list_group = 8; L.list_kind = 0;
    if (tev) CATAN_CUDA(cudaEventRecord(tev[4], env->lr_stream));
    catanb::lr_slow_kernel<<<env->lr_grid, catanb::kLrSlowThreads, sizeof(catanb::LrSmem), env->lr_stream>>>(L);
    CATAN_CUDA(cudaGetLastError());
    if (tev) CATAN_CUDA(cudaEventRecord(tev[5], env->lr_stream));
    catanb::encode_kernel<catanb::MODE_STEP, SAMPLE, true, catanb::ROLE_BOTH><<<env->sm_count, catanb::kEncThreads, sizeof(catanb::EncSmem), env->lr_stream>>>(L);
    CATAN_CUDA(cudaGetLastError());
    if (tev) CATAN_CUDA(cudaEventRecord(tev[6], env->lr_stream));
    CATAN_CUDA(cudaEventRecord(env->ev_join, env->lr_stream));
  }
  {   // the games that ended: done / reward -> reset -> encode of the new game, ONE game per block (the reset is serial)
    EnvParams L = P;
    L.list_queue = env->rs_queue; L.list_stage = env->stage_rs; L.list_count = &env->lr_ctl->rs_count; L.list_group = 1; L.list_kind = 1;
    catanb::encode_kernel<catanb::MODE_STEP, SAMPLE, true, catanb::ROLE_BOTH><<<env->sm_count * 2, catanb::kEncThreads, sizeof(catanb::EncSmem), env->rs_stream>>>(L);
    CATAN_CUDA(cudaGetLastError());
    if (tev) CATAN_CUDA(cudaEventRecord(tev[7], env->rs_stream));
    CATAN_CUDA(cudaEventRecord(env->ev_join_rs, env->rs_stream));
  }
  if (launch_encode<catanb::MODE_STEP, SAMPLE>(env, P, 0, env->n, stream, tev)) return -1;
  if (tev) CATAN_CUDA(cudaEventRecord(tev[2], stream));
  CATAN_CUDA(cudaStreamWaitEvent(stream, env->ev_join, 0));
  CATAN_CUDA(cudaStreamWaitEvent(stream, env->ev_join_rs, 0));
  return 0;
}

static void drop_graphs(catan_env* env) {
  for (auto& g : env->graphs) cudaGraphExecDestroy(g.exec);
  env->graphs.clear();
}

// Replay (capturing it first if need be) the work `record(capture stream)` issues, keyed by (kind, p0..p3).  Returns 1 when the
// caller has to issue the work directly (graphs off, timing hooks on, or a capture is already in progress on its stream).
template <class Record>
static int replay_step_graph(catan_env* env, int kind, const void* p0, const void* p1, const void* p2, const void* p3, cudaStream_t stream,
                             Record record) {
  if (!env->use_graphs || env->timing) return 1;
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(stream, &st) != cudaSuccess) { cudaGetLastError(); return 1; }
  if (st != cudaStreamCaptureStatusNone) return 1;                   // (the caller is building its own graph: become part of it)
  for (auto& g : env->graphs)
    if (g.kind == kind && g.p[0] == p0 && g.p[1] == p1 && g.p[2] == p2 && g.p[3] == p3) {
      CATAN_CUDA(cudaGraphLaunch(g.exec, stream));
      return 0;
    }
  if (!env->capture_stream) CATAN_CUDA(cudaStreamCreateWithFlags(&env->capture_stream, cudaStreamNonBlocking));
  cudaGraph_t graph = nullptr;
  CATAN_CUDA(cudaStreamBeginCapture(env->capture_stream, cudaStreamCaptureModeThreadLocal));
  const int rc = record(env->capture_stream);
  cudaError_t e = cudaStreamEndCapture(env->capture_stream, &graph);
  if (rc) { if (graph) cudaGraphDestroy(graph); return -1; }
  if (e != cudaSuccess) return cuda_fail(e, "cudaStreamEndCapture");
  cudaGraphExec_t exec = nullptr;
  e = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGraphInstantiate");
  if (env->graphs.size() >= 64) drop_graphs(env);                    // (callers use a handful of buffer sets)
  env->graphs.push_back({kind, {p0, p1, p2, p3}, exec});
  CATAN_CUDA(cudaGraphLaunch(exec, stream));
  return 0;
}

static int check_bound(const catan_env* env) {
  if (!env) return fail("null handle");
  if (!env->obs || !env->masks || !env->reward || !env->info) return fail("catan_bind has not been called");
  return 0;
}

static void free_env(catan_env* env) {
  drop_graphs(env);
  if (env->capture_stream) cudaStreamDestroy(env->capture_stream);
  if (env->lr_stream) cudaStreamDestroy(env->lr_stream);
  if (env->rs_stream) cudaStreamDestroy(env->rs_stream);
  if (env->rows_stream) cudaStreamDestroy(env->rows_stream);
  if (env->ev_fork) cudaEventDestroy(env->ev_fork);
  if (env->ev_join) cudaEventDestroy(env->ev_join);
  if (env->ev_join_rs) cudaEventDestroy(env->ev_join_rs);
  if (env->ev_join_rows) cudaEventDestroy(env->ev_join_rows);
  for (auto& slot : env->tev) for (cudaEvent_t ev : slot) if (ev) cudaEventDestroy(ev);
  cudaFree(env->recs); cudaFree(env->stage); cudaFree(env->err_flags); cudaFree(env->side); cudaFree(env->lr_slow_queue); cudaFree(env->lr_ctl);
  cudaFree(env->rs_queue); cudaFree(env->stage_rs);
  cudaFree(env->actions_stage); cudaFree(env->actions_u8_stage);
  delete env;
}

extern "C" {

int catan_abi_version(void) { return 1; }
int catan_obs_stride(void) { return CATAN_OBS_STRIDE; }
int catan_mask_stride(void) { return CATAN_MASK_STRIDE; }
int catan_info_stride(void) { return CATAN_INFO_STRIDE; }
int catan_action_words(void) { return CATAN_ACTION_WORDS; }
int catan_state_words(void) { return CATAN_STATE_WORDS; }
int catan_record_bytes(void) { return static_cast<int>(sizeof(GameRec)); }
const char* catan_last_error(void) { return g_last_error.c_str(); }

void catan_default_config(catan_config_t* c) {
  c->max_actions_per_turn = -1;
  c->max_proposed_trades_per_turn = 4;
  c->validate_actions = 1;
  c->dense_reward = 0;
  c->auto_reset = 1;
  c->win_reward = 500.0f;
  c->reward_annealing_factor = 1.0f;
}

int catan_create(int n_envs, int device, uint64_t seed, uint64_t first_env_id, const catan_config_t* cfg, catan_env_t** out) {
  if (!out) return fail("out is null");
  *out = nullptr;
  if (n_envs <= 0) return fail("n_envs must be positive");
  if (n_envs >= (1 << 28)) return fail("n_envs must be below 2^28");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) return fail("no CUDA device available: this library has no CPU path");
  if (device < 0 || device >= count) return fail("bad device index");
  catanb::DeviceScope device_scope_;                                 // (the caller's current device comes back on return)
  if (device_scope_.enter(device)) return fail("cannot make the device current");
  catan_env* env = new (std::nothrow) catan_env();
  if (!env) return fail("out of host memory");
  env->n = n_envs; env->device = device; env->seed = seed; env->first_env_id = first_env_id;
  if (cfg) env->cfg = *cfg; else catan_default_config(&env->cfg);
  cudaDeviceProp prop{};
  CATAN_CUDA(cudaGetDeviceProperties(&prop, device));
  env->sm_count = prop.multiProcessorCount;
  env->lr_grid = env->sm_count * catanb::kLrSlowBlocksPerSM;
  if (prop.persistingL2CacheMaxSize > 0 && prop.accessPolicyMaxWindowSize > 0 && getenv("CATAN_L2_WINDOW")) {   // opt-in: the set-aside is device-wide
    const size_t rec = CATAN_CHUNK_BYTES * ((static_cast<size_t>(n_envs) + 31) / 32);
    const size_t keep = rec < static_cast<size_t>(prop.persistingL2CacheMaxSize) ? rec : static_cast<size_t>(prop.persistingL2CacheMaxSize);
    if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, keep) == cudaSuccess)
      env->window_bytes = rec < static_cast<size_t>(prop.accessPolicyMaxWindowSize) ? rec : static_cast<size_t>(prop.accessPolicyMaxWindowSize);
    else cudaGetLastError();
  }   // search blocks stride over the queue; idle blocks exit at once
  env->rec_bytes = CATAN_CHUNK_BYTES * ((static_cast<size_t>(n_envs) + 31) / 32);
  const size_t n = static_cast<size_t>(n_envs);
  e = cudaMalloc(&env->recs, env->rec_bytes);
  if (e == cudaSuccess) e = cudaMemset(env->recs, 0, env->rec_bytes);
  if (e == cudaSuccess) e = cudaMalloc(&env->stage, env->rec_bytes);
  if (e == cudaSuccess) e = cudaMemset(env->stage, 0, env->rec_bytes);
  if (e == cudaSuccess) e = cudaMalloc(&env->err_flags, sizeof(uint32_t) * n);
  if (e == cudaSuccess) e = cudaMemset(env->err_flags, 0, sizeof(uint32_t) * n);
  if (e == cudaSuccess) e = cudaMalloc(&env->side, sizeof(uint32_t) * n);
  if (e == cudaSuccess) e = cudaMemset(env->side, 0, sizeof(uint32_t) * n);
  if (e == cudaSuccess) e = cudaMalloc(&env->lr_slow_queue, sizeof(uint64_t) * n);
  if (e == cudaSuccess) e = cudaMalloc(&env->rs_queue, sizeof(uint64_t) * n);
  if (e == cudaSuccess) e = cudaMalloc(&env->stage_rs, env->rec_bytes);
  if (e == cudaSuccess) e = cudaMemset(env->stage_rs, 0, env->rec_bytes);
  if (e == cudaSuccess) e = cudaMalloc(&env->lr_ctl, sizeof(catanb::LrCtl));
  if (e == cudaSuccess) e = cudaMemset(env->lr_ctl, 0, sizeof(catanb::LrCtl));
  if (e == cudaSuccess) e = cudaMalloc(&env->actions_stage, sizeof(int32_t) * CATAN_ACTION_WORDS * n);
  if (e == cudaSuccess) {
    int lo = 0, hi = 0;                                  // (greatest priority is the numerically lowest value)
    e = cudaDeviceGetStreamPriorityRange(&lo, &hi);
    if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&env->lr_stream, cudaStreamNonBlocking, hi);
    if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&env->rs_stream, cudaStreamNonBlocking, hi);
    if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&env->rows_stream, cudaStreamNonBlocking, lo);
  }
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&env->ev_fork, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&env->ev_join, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&env->ev_join_rs, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&env->ev_join_rows, cudaEventDisableTiming);
  {
    const int enc_bytes = static_cast<int>(sizeof(catanb::EncSmem));   // > 48 KB: opt in, per instantiation
    using namespace catanb;
#define CATAN_ENC_ATTR(K_, BYTES_)                                                                                         \
    if (e == cudaSuccess) e = cudaFuncSetAttribute(K_, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(BYTES_)); \
    if (e == cudaSuccess) e = cudaFuncSetAttribute(K_, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    CATAN_ENC_ATTR((encode_kernel<MODE_STEP, false, false, ROLE_ROWS>), enc_smem_bytes(ROLE_ROWS) + 32768)
    CATAN_ENC_ATTR((encode_kernel<MODE_STEP, true, false, ROLE_ROWS>), enc_smem_bytes(ROLE_ROWS) + 32768)
    CATAN_ENC_ATTR((encode_kernel<MODE_STEP, false, false, ROLE_MASKS>), enc_smem_bytes(ROLE_MASKS) + 32768)
    CATAN_ENC_ATTR((encode_kernel<MODE_STEP, true, false, ROLE_MASKS>), enc_smem_bytes(ROLE_MASKS) + 32768)
    CATAN_ENC_ATTR((encode_kernel<MODE_STEP, false, true, ROLE_BOTH>), enc_bytes)
    CATAN_ENC_ATTR((encode_kernel<MODE_STEP, true, true, ROLE_BOTH>), enc_bytes)
    CATAN_ENC_ATTR((encode_kernel<MODE_RESET, false, false, ROLE_BOTH>), enc_bytes)
    CATAN_ENC_ATTR((encode_kernel<MODE_REFRESH, false, false, ROLE_BOTH>), enc_bytes)
#undef CATAN_ENC_ATTR
  }
  if (e == cudaSuccess) e = cudaFuncSetAttribute(catanb::transition_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  // DIRECT: 16 blocks x 11.4 KB = 182 KB of shared memory per SM; the rest of the 256 KB is L1 for the cold fields touched in place
  // and for the stack of the rule functions
  if (e == cudaSuccess) e = cudaFuncSetAttribute(catanb::transition_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                                 getenv("CATAN_TRANS_CARVEOUT") ? atoi(getenv("CATAN_TRANS_CARVEOUT")) : 80);
  env->enc_pad_bytes = getenv("CATAN_ENC_PAD_KB") ? static_cast<size_t>(atoi(getenv("CATAN_ENC_PAD_KB"))) * 1024 : 0;
  env->rows_beside = getenv("CATAN_ROWS_BESIDE") != nullptr && atoi(getenv("CATAN_ROWS_BESIDE")) != 0;   // (measured: 0.300 ms per step beside, 0.279 in front)
  env->trans_direct = !(getenv("CATAN_TRANS_DIRECT") != nullptr && atoi(getenv("CATAN_TRANS_DIRECT")) == 0);   // (0: stage whole chunks)
  if (e == cudaSuccess) e = cudaFuncSetAttribute(catanb::lr_slow_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(catanb::LrSmem)));
  if (e != cudaSuccess) {
    free_env(env);
    return cuda_fail(e, "catan_create");
  }
  *out = env;
  return 0;
}

int catan_destroy(catan_env_t* env) {
  if (!env) return 0;
  free_env(env);
  return 0;
}

int catan_num_envs(const catan_env_t* env) { return env ? env->n : 0; }

int catan_set_config(catan_env_t* env, const catan_config_t* cfg) {
  if (!env || !cfg) return fail("null argument");
  env->cfg = *cfg;
  drop_graphs(env);                                                  // (the config is baked into the captured launches)
  return 0;
}

int catan_bind(catan_env_t* env, uint8_t* obs_dev, uint8_t* masks_dev, float* reward_dev, uint8_t* info_dev) {
  if (!env) return fail("null handle");
  if (!obs_dev || !masks_dev || !reward_dev || !info_dev) return fail("null output buffer");
  if ((reinterpret_cast<uintptr_t>(obs_dev) | reinterpret_cast<uintptr_t>(masks_dev) | reinterpret_cast<uintptr_t>(reward_dev) |
       reinterpret_cast<uintptr_t>(info_dev)) & 15)
    return fail("output buffers must be 16-byte aligned");
  if (reinterpret_cast<uintptr_t>(obs_dev) & 31) return fail("the observation buffer must be 32-byte aligned (rows are written with 256-bit stores)");
  env->obs = obs_dev; env->masks = masks_dev; env->reward = reward_dev; env->info = info_dev;
  drop_graphs(env);
  return 0;
}

int catan_reset(catan_env_t* env, const uint8_t* reset_mask_dev, void* stream) {
  if (check_bound(env)) return -1;
  CATAN_ON_DEVICE_OF(env);
  EnvParams P = make_params(env);
  P.env_mask = reset_mask_dev;
  return launch_encode<catanb::MODE_RESET, false>(env, P, 0, env->n, static_cast<cudaStream_t>(stream));
}

int catan_step(catan_env_t* env, const int32_t* actions_dev, void* stream) { return catan_step_masked(env, actions_dev, nullptr, stream); }

int catan_step_masked(catan_env_t* env, const int32_t* actions_dev, const uint8_t* step_mask_dev, void* stream) {
  if (check_bound(env)) return -1;
  if (!actions_dev) return fail("actions_dev is null");
  CATAN_ON_DEVICE_OF(env);
  EnvParams P = make_params(env);
  P.actions = actions_dev;
  P.env_mask = step_mask_dev;
  const int g = replay_step_graph(env, 1, actions_dev, step_mask_dev, nullptr, nullptr, static_cast<cudaStream_t>(stream),
                                  [&](cudaStream_t s) { return launch_step<false>(env, P, s); });
  return g <= 0 ? g : launch_step<false>(env, P, static_cast<cudaStream_t>(stream));
}

int catan_step_sample(catan_env_t* env, int32_t* actions_io_dev, void* stream) {
  if (check_bound(env)) return -1;
  if (!actions_io_dev) return fail("actions_io_dev is null");
  CATAN_ON_DEVICE_OF(env);
  EnvParams P = make_params(env);
  P.actions = actions_io_dev;
  P.actions_out = actions_io_dev;
  const int g = replay_step_graph(env, 2, actions_io_dev, nullptr, nullptr, nullptr, static_cast<cudaStream_t>(stream),
                                  [&](cudaStream_t s) { return launch_step<true>(env, P, s); });
  return g <= 0 ? g : launch_step<true>(env, P, static_cast<cudaStream_t>(stream));
}

int catan_sample_random(catan_env_t* env, int32_t* actions_out_dev, void* stream) {
  if (check_bound(env)) return -1;
  if (!actions_out_dev) return fail("actions_out_dev is null");
  CATAN_ON_DEVICE_OF(env);
  const int blocks = (env->n + catanb::kSampleThreads - 1) / catanb::kSampleThreads;
  catanb::sample_kernel<<<blocks, catanb::kSampleThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      env->recs, env->n, env->seed, env->first_env_id, env->masks, env->obs, actions_out_dev);
  CATAN_CUDA(cudaGetLastError());
  return 0;
}

static int copy_outputs_to_host(catan_env_t* env, uint8_t* obs_host, uint8_t* masks_host, float* reward_host, uint8_t* info_host,
                                cudaStream_t s, bool sync = true) {
  const size_t n = static_cast<size_t>(env->n);
  if (obs_host) CATAN_CUDA(cudaMemcpyAsync(obs_host, env->obs, n * CATAN_OBS_STRIDE, cudaMemcpyDeviceToHost, s));
  if (masks_host) CATAN_CUDA(cudaMemcpyAsync(masks_host, env->masks, n * CATAN_MASK_STRIDE, cudaMemcpyDeviceToHost, s));
  if (reward_host) CATAN_CUDA(cudaMemcpyAsync(reward_host, env->reward, n * 4 * sizeof(float), cudaMemcpyDeviceToHost, s));
  if (info_host) CATAN_CUDA(cudaMemcpyAsync(info_host, env->info, n * CATAN_INFO_STRIDE, cudaMemcpyDeviceToHost, s));
  if (sync) CATAN_CUDA(cudaStreamSynchronize(s));
  return 0;
}

// The same call without the final synchronisation: the host buffers (pinned, or the copies serialise) hold the step's result
// once `stream` has been synchronised.  Two handles on two streams, each owning half of the games, overlap one half's PCIe
// copies with the other half's kernels.
int catan_step_host_async(catan_env_t* env, const int32_t* actions_host, uint8_t* obs_host, uint8_t* masks_host, float* reward_host,
                          uint8_t* info_host, void* stream) {
  if (check_bound(env)) return -1;
  if (!actions_host) return fail("actions_host is null");
  CATAN_ON_DEVICE_OF(env);
  auto issue = [&](cudaStream_t s) -> int {
    CATAN_CUDA(cudaMemcpyAsync(env->actions_stage, actions_host, sizeof(int32_t) * CATAN_ACTION_WORDS * static_cast<size_t>(env->n),
                               cudaMemcpyHostToDevice, s));
    EnvParams P = make_params(env);
    P.actions = env->actions_stage;
    if (launch_step<false>(env, P, s)) return -1;
    return copy_outputs_to_host(env, obs_host, masks_host, reward_host, info_host, s, false);
  };
  // (the graph is keyed by the host pointers: pinned buffers that the caller reuses every tick)
  const void* k3 = obs_host ? static_cast<const void*>(obs_host) : static_cast<const void*>(masks_host);
  const int g = replay_step_graph(env, 3, actions_host, reward_host, info_host, k3, static_cast<cudaStream_t>(stream), issue);
  return g <= 0 ? g : issue(static_cast<cudaStream_t>(stream));
}

// catan_step_sample with host buffers, not synchronised: actions_io_host (pinned) is copied in, applied, and overwritten with
// the next random-legal action of every env; reward / info rows as in catan_step_host_async.
int catan_step_sample_host_async(catan_env_t* env, int32_t* actions_io_host, float* reward_host, uint8_t* info_host, void* stream) {
  if (check_bound(env)) return -1;
  if (!actions_io_host) return fail("actions_io_host is null");
  CATAN_ON_DEVICE_OF(env);
  const size_t bytes = sizeof(int32_t) * CATAN_ACTION_WORDS * static_cast<size_t>(env->n);
  auto issue = [&](cudaStream_t s) -> int {
    CATAN_CUDA(cudaMemcpyAsync(env->actions_stage, actions_io_host, bytes, cudaMemcpyHostToDevice, s));
    EnvParams P = make_params(env);
    P.actions = env->actions_stage;
    P.actions_out = env->actions_stage;
    if (launch_step<true>(env, P, s)) return -1;
    CATAN_CUDA(cudaMemcpyAsync(actions_io_host, env->actions_stage, bytes, cudaMemcpyDeviceToHost, s));
    return copy_outputs_to_host(env, nullptr, nullptr, reward_host, info_host, s, false);
  };
  const int g = replay_step_graph(env, 4, actions_io_host, reward_host, info_host, nullptr, static_cast<cudaStream_t>(stream), issue);
  return g <= 0 ? g : issue(static_cast<cudaStream_t>(stream));
}

int catan_step_sample_host_async_u8(catan_env_t* env, uint8_t* actions_io_host_u8, float* reward_host, uint8_t* info_host, void* stream) {
  if (check_bound(env)) return -1;
  if (!actions_io_host_u8) return fail("actions_io_host_u8 is null");
  CATAN_ON_DEVICE_OF(env);
  const int n_words = CATAN_ACTION_WORDS * env->n;
  static_assert(CATAN_ACTION_WORDS % 4 == 0, "the pack / unpack kernels move four words per thread");
  if (!env->actions_u8_stage) CATAN_CUDA(cudaMalloc(&env->actions_u8_stage, static_cast<size_t>(n_words)));
  auto issue = [&](cudaStream_t s) -> int {
    CATAN_CUDA(cudaMemcpyAsync(env->actions_u8_stage, actions_io_host_u8, static_cast<size_t>(n_words), cudaMemcpyHostToDevice, s));
    catanb::actions_unpack_kernel<<<(n_words / 4 + 255) / 256, 256, 0, s>>>(env->actions_u8_stage, env->actions_stage, n_words);
    CATAN_CUDA(cudaGetLastError());
    EnvParams P = make_params(env);
    P.actions = env->actions_stage;
    P.actions_out = env->actions_stage;
    if (launch_step<true>(env, P, s)) return -1;
    catanb::actions_pack_kernel<<<(n_words / 4 + 255) / 256, 256, 0, s>>>(env->actions_stage, env->actions_u8_stage, n_words);
    CATAN_CUDA(cudaGetLastError());
    CATAN_CUDA(cudaMemcpyAsync(actions_io_host_u8, env->actions_u8_stage, static_cast<size_t>(n_words), cudaMemcpyDeviceToHost, s));
    return copy_outputs_to_host(env, nullptr, nullptr, reward_host, info_host, s, false);
  };
  const int g = replay_step_graph(env, 5, actions_io_host_u8, reward_host, info_host, nullptr, static_cast<cudaStream_t>(stream), issue);
  return g <= 0 ? g : issue(static_cast<cudaStream_t>(stream));
}

int catan_step_sample_host_groups(catan_env_t* const* envs, int n_groups, void* const* actions_io_host, int action_format, float* const* reward_host,
                                  uint8_t* const* info_host, void* const* streams, int rounds, long long* done_seen) {
  if (action_format != CATAN_ACTIONS_I32 && action_format != CATAN_ACTIONS_U8) return fail("catan_step_sample_host_groups: bad action format");
  if (!envs || !actions_io_host || !reward_host || !info_host || !streams || n_groups <= 0 || rounds < 0) return fail("catan_step_sample_host_groups: bad argument");
  long long seen = 0;
  for (int r = 0; r < rounds; ++r) {
    for (int g = 0; g < n_groups; ++g) {
      if (check_bound(envs[g])) return -1;
      {
        CATAN_ON_DEVICE_OF(envs[g]);
        CATAN_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(streams[g])));
      }
      if (info_host[g]) {
        const uint8_t* info = info_host[g];
        const int n = envs[g]->n;
        for (int i = 0; i < n; ++i) seen += info[static_cast<size_t>(i) * CATAN_INFO_STRIDE + CATAN_INFO_DONE];
      }
      if (action_format == CATAN_ACTIONS_U8 ? catan_step_sample_host_async_u8(envs[g], static_cast<uint8_t*>(actions_io_host[g]), reward_host[g], info_host[g], streams[g])
                                            : catan_step_sample_host_async(envs[g], static_cast<int32_t*>(actions_io_host[g]), reward_host[g], info_host[g], streams[g]))
        return -1;
    }
  }
  if (done_seen) *done_seen += seen;
  return 0;
}

int catan_step_host(catan_env_t* env, const int32_t* actions_host, uint8_t* obs_host, uint8_t* masks_host, float* reward_host,
                    uint8_t* info_host, void* stream) {
  if (check_bound(env)) return -1;
  if (!actions_host) return fail("actions_host is null");
  CATAN_ON_DEVICE_OF(env);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CATAN_CUDA(cudaMemcpyAsync(env->actions_stage, actions_host, sizeof(int32_t) * CATAN_ACTION_WORDS * static_cast<size_t>(env->n),
                             cudaMemcpyHostToDevice, s));
  if (catan_step(env, env->actions_stage, stream)) return -1;
  return copy_outputs_to_host(env, obs_host, masks_host, reward_host, info_host, s);
}

int catan_reset_host(catan_env_t* env, uint8_t* obs_host, uint8_t* masks_host, uint8_t* info_host, void* stream) {
  if (catan_reset(env, nullptr, stream)) return -1;
  return copy_outputs_to_host(env, obs_host, masks_host, nullptr, info_host, static_cast<cudaStream_t>(stream));
}

// the chunks [first / 32, (first + count - 1) / 32] of the record array <-> host
static int chunk_range(const catan_env* env, int first, int count, size_t& off, size_t& bytes) {
  const size_t c0 = static_cast<size_t>(first) / 32, c1 = static_cast<size_t>(first + count - 1) / 32;
  off = c0 * CATAN_CHUNK_BYTES;
  bytes = (c1 - c0 + 1) * CATAN_CHUNK_BYTES;
  return off + bytes <= env->rec_bytes ? 0 : fail("env range out of bounds");
}

int catan_export_state(catan_env_t* env, int first, int count, int16_t* states_host) {
  if (!env || !states_host) return fail("null argument");
  if (first < 0 || count < 0 || first + count > env->n) return fail("env range out of bounds");
  if (count == 0) return 0;
  CATAN_ON_DEVICE_OF(env);
  size_t off, bytes;
  if (chunk_range(env, first, count, off, bytes)) return -1;
  std::vector<uint8_t> tmp(bytes);
  CATAN_CUDA(cudaDeviceSynchronize());
  CATAN_CUDA(cudaMemcpy(tmp.data(), env->recs + off, bytes, cudaMemcpyDeviceToHost));
  catan_state_t* out = reinterpret_cast<catan_state_t*>(states_host);
  const int base = first & ~31;
  for (int i = 0; i < count; ++i) {
    const int e = first + i - base;
    GameRec rec;
    catanb::chunk_get(tmp.data() + static_cast<size_t>(e / 32) * CATAN_CHUNK_BYTES, e % 32, 32, rec);
    catanb::rec_to_state(rec, out[i]);
  }
  return 0;
}

int catan_import_state(catan_env_t* env, int first, int count, const int16_t* states_host) {
  if (check_bound(env)) return -1;
  if (!states_host) return fail("null argument");
  if (first < 0 || count < 0 || first + count > env->n) return fail("env range out of bounds");
  if (count == 0) return 0;
  CATAN_ON_DEVICE_OF(env);
  size_t off, bytes;
  if (chunk_range(env, first, count, off, bytes)) return -1;
  std::vector<uint8_t> tmp(bytes);
  CATAN_CUDA(cudaDeviceSynchronize());
  CATAN_CUDA(cudaMemcpy(tmp.data(), env->recs + off, bytes, cudaMemcpyDeviceToHost));
  const catan_state_t* in = reinterpret_cast<const catan_state_t*>(states_host);
  const int base = first & ~31;
  for (int i = 0; i < count; ++i) {
    const int e = first + i - base;
    uint8_t* chunk = tmp.data() + static_cast<size_t>(e / 32) * CATAN_CHUNK_BYTES;
    GameRec rec;
    catanb::chunk_get(chunk, e % 32, 32, rec);   // keeps what the canonical state does not carry
    catanb::state_to_rec(in[i], rec);
    catanb::chunk_put(chunk, e % 32, 32, rec);
  }
  CATAN_CUDA(cudaMemcpy(env->recs + off, tmp.data(), bytes, cudaMemcpyHostToDevice));
  EnvParams P = make_params(env);
  if (launch_encode<catanb::MODE_REFRESH, false>(env, P, first, count, nullptr)) return -1;
  CATAN_CUDA(cudaDeviceSynchronize());
  return 0;
}

int catan_randomise_uncertainty(catan_env_t* env, const uint8_t* controlling_pid_dev, int max_attempts, void* stream) {
  if (check_bound(env)) return -1;
  if (!controlling_pid_dev || max_attempts <= 0) return fail("catan_randomise_uncertainty: bad argument");
  CATAN_ON_DEVICE_OF(env);
  EnvParams P = make_params(env);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  catanb::randomise_kernel<<<(env->n + 127) / 128, 128, 0, s>>>(P, controlling_pid_dev, max_attempts);
  CATAN_CUDA(cudaGetLastError());
  return launch_encode<catanb::MODE_REFRESH, false>(env, P, 0, env->n, s);   // observations / masks of the re-dealt games
}

int catan_read_lr_stats(catan_env_t* env, unsigned long long* out_host) {
  if (!env || !out_host) return fail("null argument");
  CATAN_ON_DEVICE_OF(env);
  catanb::LrCtl last;
  CATAN_CUDA(cudaDeviceSynchronize());
  CATAN_CUDA(cudaMemcpy(&last, env->lr_ctl, sizeof(last), cudaMemcpyDeviceToHost));
  out_host[0] = last.total; out_host[1] = last.slow_total;
  out_host[2] = last.dbg[4]; out_host[3] = last.dbg[5];
  for (int i = 0; i < 4; ++i) out_host[4 + i] = last.dbg[i];
  return 0;
}

int catan_read_lr_histograms(catan_env_t* env, unsigned long long* out_host) {
  if (!env || !out_host) return fail("null argument");
  CATAN_ON_DEVICE_OF(env);
  catanb::LrCtl last;
  CATAN_CUDA(cudaDeviceSynchronize());
  CATAN_CUDA(cudaMemcpy(&last, env->lr_ctl, sizeof(last), cudaMemcpyDeviceToHost));
  memcpy(out_host, last.hist, sizeof(last.hist));
  return 0;
}

int catan_set_graphs(catan_env_t* env, int enable) {
  if (!env) return fail("null handle");
  CATAN_ON_DEVICE_OF(env);
  env->use_graphs = enable != 0;
  if (!enable) { CATAN_CUDA(cudaDeviceSynchronize()); drop_graphs(env); }
  return 0;
}

int catan_set_timing(catan_env_t* env, int enable) {
  if (!env) return fail("null handle");
  CATAN_ON_DEVICE_OF(env);
  CATAN_CUDA(cudaDeviceSynchronize());
  if (enable)
    for (auto& slot : env->tev) for (cudaEvent_t& ev : slot) if (!ev) CATAN_CUDA(cudaEventCreate(&ev));
  env->timing = enable != 0; env->timed = 0; for (double& x : env->t_ms) x = 0.0; env->t_n = 0;
  return 0;
}

int catan_read_timing(catan_env_t* env, double* out_host) {
  if (!env || !out_host) return fail("null argument");
  CATAN_ON_DEVICE_OF(env);
  const unsigned long long pending = env->timed < 32 ? env->timed : 32;
  for (unsigned long long k = env->timed - pending; k < env->timed; ++k) if (retire_timed_step(env, env->tev[k % 32])) return -1;
  env->timed = 0;                                                    // (the ring is empty again)
  out_host[0] = static_cast<double>(env->t_n); out_host[1] = env->t_ms[0]; out_host[2] = env->t_ms[1]; out_host[3] = env->t_ms[2];
  for (int k = 3; k < 8; ++k) out_host[1 + k] = env->t_ms[k];
  return 0;
}

#ifdef CATAN_PROFILE_PHASES
int catan_debug_read_phases(unsigned long long* out64_host, int clear) {
  if (cudaMemcpyFromSymbol(out64_host, catanb::d_phase, sizeof(unsigned long long) * 64) != cudaSuccess) return fail("catan_debug_read_phases");
  if (clear) { unsigned long long z[64] = {0}; cudaMemcpyToSymbol(catanb::d_phase, z, sizeof(z)); }
  return 0;
}
#endif

int catan_read_err_flags(catan_env_t* env, uint32_t* flags_host, int clear) {
  if (!env || !flags_host) return fail("null argument");
  CATAN_ON_DEVICE_OF(env);
  CATAN_CUDA(cudaDeviceSynchronize());
  CATAN_CUDA(cudaMemcpy(flags_host, env->err_flags, sizeof(uint32_t) * static_cast<size_t>(env->n), cudaMemcpyDeviceToHost));
  if (clear) CATAN_CUDA(cudaMemset(env->err_flags, 0, sizeof(uint32_t) * static_cast<size_t>(env->n)));
  return 0;
}

}  // extern "C"
