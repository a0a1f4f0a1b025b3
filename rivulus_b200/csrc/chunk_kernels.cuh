// chunk_kernels.cuh — single-pass Filter + Select for queries whose PREDICATE column is also PROJECTED
// (`filter(k > T)` with no Select in front of it projects every column, k included; BASELINE configs[4] is this shape).
//
// The two-pass plan reads such a column twice: once in the predicate scan, once more in the compaction pass (B_alg counts it
// once; at 50 % that second read is a quarter of all traffic of configs[4]).  Here the predicate values are read from HBM
// exactly ONCE: a persistent CTA keeps a CHUNK of them (2 tiles = 4096 rows = 32 KB) in shared memory from the moment the
// predicate is evaluated until its survivors have been written out.  Chunks are handed out by a ticket, in row order.
//
// Global output order.  Every chunk publishes its survivor count as a descriptor the moment its predicate pass ends; ONE warp of
// the whole grid — the PREFIX SERVER, hosted by whichever CTA drew ticket 0 (so it is resident by construction) — walks the
// descriptors in chunk order, 32 per step with the next step's loads already in flight, and turns them into inclusive prefixes
// in place.  A chunk then needs exactly one word: its own descriptor.  (First version: a 256-wide decoupled look-back per chunk,
// one per CTA at a time.  With ~450 chunks in flight every look-back re-read the same few cache lines as 147 others and waited on
// the slowest of its predecessors: 7.7 us per chunk and CTA, 2.7x slower than the two-pass plan.  The server does the same sums
// once, sequentially, at > 1000 chunks/us.)
//
// One CTA per SM, 19 warps:
//   warp 16  producer   one lane, an event loop that never blocks on one thing while another is ready: draws the next ticket,
//                       TMA-loads that chunk's predicate values into one of four chunk buffers as soon as the buffer is free, and
//                       — for every chunk whose tile counts are known — streams the OTHER projected 8-byte columns of its dense
//                       tiles through a ring of 16 KB slots (cp.async.bulk + mbarrier complete_tx)
//   warp 17  scanner    per chunk: waits for its own descriptor to become a prefix, writes the tile_info words for the bit-packed
//                       / string kernels that follow, hands the prefix to the consumers
//   warp 18  server     (only in the CTA that drew ticket 0) the grid-wide descriptor scan described above
//   warps 0-15 consumers  iteration i: (P1) evaluate the predicate of chunk i out of its buffer (lane <-> row, one ballot = one
//                       selection word) and publish its count, then (P2) compact chunk i-2: the predicate column from its chunk
//                       buffer, the other columns from the ring (dense tiles) or with 64-byte-granule gather loads (sparse tiles),
//                       survivors stored straight to their final position.  P2 lags P1 by two chunks (~4 us): the round trip
//                       count -> server -> prefix and the first column tiles are there when P2 needs them.
//
// HBM traffic: predicate column 8 B/row once, every other projected column as in the two-pass plan, survivors written once,
// plus the selection bitmap (1 bit/row) for the bit-packed and string columns.
#pragma once
#include "compact_kernels.cuh"

namespace rvl {

constexpr int kChunkTiles = 2;                           // 2048-row tiles per chunk
constexpr int kChunkRows = kChunkTiles * kTileRows;      // 4096
constexpr int kChunkWords = kChunkRows / 32;             // 128 selection words
#ifndef RVL_CHUNK_BUFS
#define RVL_CHUNK_BUFS 4
#define RVL_CHUNK_LAG 2
#endif
constexpr int kChunkBufs = RVL_CHUNK_BUFS;               // chunk buffers: being loaded / evaluated / waiting for its prefix (kChunkLag of them) / compacted
constexpr int kChunkLag = RVL_CHUNK_LAG;                 // P2 runs this many chunks behind P1.  Measured on a 500 M-row shard at 50 % (ms): (bufs, lag) =
                                                         // (4, 2) 2.65, (5, 3) 2.73, (5, 2) 2.71 — neither a longer lag nor one more chunk loading
                                                         // ahead helps; the smaller ring (3 slots instead of 5) costs a little
static_assert(kChunkBufs >= kChunkLag + 2, "one buffer being loaded, one evaluated, kChunkLag waiting for their prefix");
constexpr int kChunkCW = 16;                             // consumer warps
constexpr int kChunkThreads = (kChunkCW + 3) * 32;       // + producer, scanner, server
constexpr int kChunkMaxSlots = 5;
#ifndef RVL_CHUNK_SERVER_PER
#define RVL_CHUNK_SERVER_PER 4
#endif
constexpr int kServerPer = RVL_CHUNK_SERVER_PER;         // descriptors per lane and step of the prefix server

struct ChunkParams {
    int64_t n_rows;
    int64_t n_chunks;
    const uint64_t* pred_values;       // row 0 of the view (16-byte aligned: checked by the host)
    int64_t lit_bits;
    uint64_t range_lo, range_span;
    uint32_t range_neg, truth, keep_null;
    int32_t pred_col;                  // index into col8[] of the predicate column
    BitSrc pred_valid;
    int32_t n_col8, n_slots;
    uint32_t sparse_max, debug;        // debug (RVL_CHUNK_DEBUG, timing experiments only): 1 = skip the look-back (WRONG output order)
    uint32_t producer_nap, pad_nap;    // ns the producer lane sleeps after an event-loop pass that found nothing to do (0 = spin)
    Col8 col8[kMaxCol8];
    const unsigned long long* base_in; // rows emitted by earlier launches of the same query (streaming) or nullptr
    uint32_t* sel_out;                 // row-order selection words, whole tiles
    uint64_t* tile_info;               // per tile: (global exclusive output index << 12) | survivors
    uint64_t* status;                  // one look-back descriptor per chunk, zeroed
    uint32_t* ticket;                  // chunk dispenser, zeroed
    unsigned long long* total_out;     // base + survivors
    unsigned long long* debug_words;   // pinned host words (RVL_CHUNK_DEBUG=2)
};

__device__ __forceinline__ void chunk_stuck(const ChunkParams& p, uint32_t site, int64_t a, int64_t b) {
    if (p.debug_words != nullptr) {
        unsigned long long* w = p.debug_words + 8 * (site & 15u);   // one record per wait site: the first reporter wins
        if (atomicCAS(w + 5, 0ull, 1ull) == 0ull) {
            w[0] = site; w[1] = blockIdx.x; w[2] = threadIdx.x >> 5; w[3] = (unsigned long long)a; w[4] = (unsigned long long)b;
        }
        atomicAdd(w + 6, 1ull);
        __threadfence_system();
    }
    __nanosleep(20000000);   // let the other stuck warps report before the trap tears the context down
    __trap();
}

struct __align__(128) ChunkSmem {
    uint64_t kbuf[kChunkBufs][kChunkRows];       // 128 KB
    uint64_t kfull[kChunkBufs], kfree[kChunkBufs], p1done[kChunkBufs], pready[kChunkBufs];
    uint64_t full[kChunkMaxSlots], empty[kChunkMaxSlots];
    uint32_t sel[kChunkBufs][kChunkWords];
    uint32_t wcnt[kChunkBufs][kChunkCW];
    uint32_t ctot[kChunkBufs], arrived[kChunkBufs];   // P1: running survivor total / consumer warps done — the last one publishes the aggregate
    int64_t chunk_id[kChunkBufs];
    uint64_t prefix[kChunkBufs];
    uint32_t server_role;                        // 0 = undecided, 1 = this CTA drew ticket 0 (hosts the prefix server), 2 = it did not
    uint32_t pad[3];
    // ring slots follow (dynamic): n_slots x 16 KB, 128-byte aligned
};

// RVL_CHUNK_DEBUG=2 (debugging only): every wait gives up after ~2^24 polls, leaves {site, block, warp, a, b} in the pinned debug
// words and traps, so a protocol bug shows where it is stuck instead of hanging the device.
__device__ __forceinline__ void chunk_stuck(const ChunkParams& p, uint32_t site, int64_t a, int64_t b);
#define CHUNK_WAIT(cond, site, a, b)                                        \
    do {                                                                    \
        uint32_t spins__ = 0;                                               \
        while (!(cond)) {                                                   \
            if (p.debug == 2u && ++spins__ > (1u << 24)) chunk_stuck(p, site, a, b); \
        }                                                                   \
    } while (0)

template <int PRED>
__device__ __forceinline__ bool chunk_keep(const ChunkParams& p, uint64_t v) {
    if (PRED == kPredI64) return ((v - p.range_lo) <= p.range_span) != (p.range_neg != 0u);
    return (p.truth & cmp_code<PRED>(v, p.lit_bits)) != 0u;
}

template <int PRED>
__global__ void __launch_bounds__(kChunkThreads, 1) chunk_filter_kernel(const __grid_constant__ ChunkParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    ChunkSmem& sm = *reinterpret_cast<ChunkSmem*>(smem_raw);
    uint64_t* const slots = reinterpret_cast<uint64_t*>(smem_raw + ((sizeof(ChunkSmem) + 127) & ~size_t(127)));
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t base0 = p.base_in != nullptr ? (uint64_t)*p.base_in : 0ull;

    if (tid == 0) {
        for (int b = 0; b < kChunkBufs; ++b) {
            mbar_init(&sm.kfull[b], 1); mbar_init(&sm.kfree[b], kChunkCW); mbar_init(&sm.p1done[b], kChunkCW); mbar_init(&sm.pready[b], 1);
        }
        for (int s = 0; s < p.n_slots; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], kChunkCW); }
        for (int b = 0; b < kChunkBufs; ++b) { sm.ctot[b] = 0u; sm.arrived[b] = 0u; }
        sm.server_role = 0u;
        mbar_init_fence();
    }
    __syncthreads();

    // a tile of a chunk goes through the ring iff it is dense, whole, and the column can be bulk-copied: the producer and the
    // consumers evaluate exactly this, from the same shared counts
    auto tile_rows_left = [&](int64_t chunk, int t) -> int64_t { return p.n_rows - (chunk * kChunkRows + (int64_t)t * kTileRows); };

    if (warp == kChunkCW) {
        // ------------------------------------------------------------------------------------------------ producer
        if (lane != 0) return;
        int slot = 0;
        uint32_t round = 0;
        int64_t next_k = 0;          // next iteration whose predicate values are to be loaded
        int64_t next_r = 0;          // iteration whose other columns are being streamed ...
        int r_tile = 0, r_col = 0;   // ... and the (tile, column) the next ring slot is for
        bool r_open = false;         // next_r's tile counts have been seen (its P1 is done)
        bool out_of_chunks = false;
        int64_t ticket = (int64_t)atomicAdd(p.ticket, 1u);   // the ticket of iteration 0; later ones are drawn one iteration ahead
        *reinterpret_cast<volatile uint32_t*>(&sm.server_role) = ticket == 0 ? 1u : 2u;
        uint32_t idle = 0;
        // An event loop: every poll is non-blocking (mbarrier.test_wait) and every step does at most one unit of work, so the
        // predicate loads (a) are never held up behind a ring slot that only frees once the consumers have been given (a)'s data.
#pragma unroll 1
        while (true) {
            if (p.debug == 2u && ++idle > (1u << 26)) chunk_stuck(p, 12, next_k, next_r);
            // a pass that found nothing to issue yields the scheduler's issue slots to the consumer warps for a moment
            if (p.producer_nap != 0u && idle > 1u) __nanosleep(p.producer_nap);
            if (p.debug != 2u) ++idle;
            // (a) load the next chunk's predicate values as soon as its buffer is free
            if (!out_of_chunks) {
                const int b = (int)(next_k % kChunkBufs);
                if (next_k < kChunkBufs || mbar_test(&sm.kfree[b], (uint32_t)((next_k / kChunkBufs - 1) & 1))) {
                    const int64_t c = ticket;
                    const bool have = c < p.n_chunks;
                    sm.chunk_id[b] = have ? c : -1;
                    if (have && c * kChunkRows + kChunkRows <= p.n_rows) {
                        // whole chunk: one transaction count, one bulk copy per tile
                        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&sm.kfull[b])), "r"((uint32_t)(kChunkRows * 8)) : "memory");
#pragma unroll
                        for (int t = 0; t < kChunkTiles; ++t)
                            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(&sm.kbuf[b][t * kTileRows])),
                                         "l"(p.pred_values + c * kChunkRows + (int64_t)t * kTileRows), "r"((uint32_t)kSlotBytes), "r"(smem_u32(&sm.kfull[b]))
                                         : "memory");
                    } else {
                        mbar_arrive(&sm.kfull[b]);   // ragged last chunk (consumers read global memory) or no chunk left
                    }
                    if (!have) out_of_chunks = true;
                    else ticket = (int64_t)atomicAdd(p.ticket, 1u);   // in flight while the loop goes on: first used an iteration later
                    ++next_k;
                    idle = 0;
                }
            }
            // (b) one ring slot of the oldest chunk whose tile counts are known: the other projected columns of its dense tiles
            const int64_t valid_k = out_of_chunks ? next_k - 1 : next_k;   // iterations [0, valid_k) hold real chunks
            if (next_r < valid_k) {
                const int pb = (int)(next_r % kChunkBufs);
                if (!r_open && mbar_test(&sm.p1done[pb], (uint32_t)((next_r / kChunkBufs) & 1))) { r_open = true; r_tile = 0; r_col = 0; }
                if (r_open) {
                    const int64_t pc = sm.chunk_id[pb];
                    // advance the cursor to the next (tile, column) that goes through the ring
                    bool found = false;
                    while (r_tile < kChunkTiles) {
                        uint32_t cnt = 0;
#pragma unroll
                        for (int w = 0; w < kChunkCW / kChunkTiles; ++w) cnt += sm.wcnt[pb][r_tile * (kChunkCW / kChunkTiles) + w];
                        if (cnt > p.sparse_max && tile_rows_left(pc, r_tile) >= kTileRows) {
                            while (r_col < p.n_col8 && (r_col == p.pred_col || !p.col8[r_col].vec_ok)) ++r_col;
                            if (r_col < p.n_col8) { found = true; break; }
                        }
                        ++r_tile; r_col = 0;
                    }
                    if (!found) { ++next_r; r_open = false; idle = 0; }
                    else if (round == 0u || mbar_test(&sm.empty[slot], (round - 1u) & 1u)) {
                        const int64_t row0 = pc * kChunkRows + (int64_t)r_tile * kTileRows;
                        tma_load_1d(slots + (size_t)slot * kTileRows, p.col8[r_col].in + row0, kSlotBytes, &sm.full[slot]);
                        if (++slot == p.n_slots) { slot = 0; ++round; }
                        ++r_col;
                        idle = 0;
                    }
                }
            } else if (out_of_chunks) {
                break;   // every real chunk has had its columns issued
            }
        }
        return;
    }

    if (warp == kChunkCW + 1) {
        // ------------------------------------------------------------------------------------------------ scanner
#pragma unroll 1
        for (int64_t i = 0;; ++i) {
            const int b = (int)(i % kChunkBufs);
            const uint32_t par = (uint32_t)((i / kChunkBufs) & 1);
            CHUNK_WAIT(mbar_try_wait(&sm.kfull[b], par), 2, i, b);          // chunk_id[b] is visible
            const int64_t c = sm.chunk_id[b];
            if (c < 0) break;
            CHUNK_WAIT(mbar_try_wait(&sm.p1done[b], par), 3, i, c);
            uint32_t tc[kChunkTiles];
            uint32_t total = 0;
#pragma unroll
            for (int t = 0; t < kChunkTiles; ++t) {
                tc[t] = 0;
#pragma unroll
                for (int w = 0; w < kChunkCW / kChunkTiles; ++w) tc[t] += sm.wcnt[b][t * (kChunkCW / kChunkTiles) + w];
                total += tc[t];
            }
            if (lane == 0) {
                // our descriptor, published as a count by the last consumer warp of P1, comes back from the server as an inclusive prefix
                uint64_t d = ld_relaxed_gpu(p.status + c);
                CHUNK_WAIT(((d = ld_relaxed_gpu(p.status + c)) & kStatusMask) == kStatusPrefix, 4, i, c);
                const uint64_t incl = p.debug == 1u ? (uint64_t)total : (d & kValueMask);
                const uint64_t excl = base0 + incl - total;
                uint64_t run = excl;
#pragma unroll
                for (int t = 0; t < kChunkTiles; ++t) {
                    if (tile_rows_left(c, t) > 0) p.tile_info[c * kChunkTiles + t] = (run << kInfoShift) | (uint64_t)tc[t];
                    run += tc[t];
                }
                if (c == p.n_chunks - 1) *p.total_out = (unsigned long long)(excl + total);
                sm.prefix[b] = excl;
                mbar_arrive(&sm.pready[b]);
            }
            __syncwarp();
        }
        return;
    }

    if (warp == kChunkCW + 2) {
        // ------------------------------------------------------------------------------------------------ prefix server
        CHUNK_WAIT(*reinterpret_cast<volatile uint32_t*>(&sm.server_role) != 0u, 11, 0, 0);
        if (*reinterpret_cast<volatile uint32_t*>(&sm.server_role) != 1u) return;
        // descriptors in chunk order, the next step's loads issued before this step is consumed.  (An 8-step look-ahead with one
        // descriptor per lane was measured SLOWER, 2.98 vs 2.66 ms: loads issued that early mostly come back unpublished and are polled
        // again, one L2 round trip each.)
        // A window of 32 * kServerPer descriptors per step (kServerPer consecutive ones per lane).  A step costs one L2 round trip
        // whatever it covers; with 32 per step that pace — 3 815 steps for 500 M rows — WAS the kernel's duration (2.67 ms at any
        // selectivity).  The step consumes the leading run of PUBLISHED descriptors and moves the window behind it: the prefix of
        // chunk c never waits for a chunk after c.  (Waiting for a whole window can deadlock: a CTA holds at most kChunkBufs chunks,
        // and when more of its tickets than that fall into one window, the later ones are only loaded once the earlier ones have
        // their prefix.  Seen as a hang at 0.1 % selectivity with 128-descriptor windows.)
        uint64_t running = 0;
        int64_t base = 0;
        uint32_t spins = 0;
#pragma unroll 1
        while (base < p.n_chunks) {
            const int64_t idx0 = base + (int64_t)lane * kServerPer;
            uint64_t d[kServerPer];
#pragma unroll
            for (int k = 0; k < kServerPer; ++k) d[k] = idx0 + k < p.n_chunks ? ld_relaxed_gpu(p.status + idx0 + k) : kStatusAggregate;
            // leading published descriptors of this lane, then of the window
            int mine_ready = 0;
#pragma unroll
            for (int k = 0; k < kServerPer; ++k) if (mine_ready == k && (d[k] & kStatusMask) != 0ull) mine_ready = k + 1;
            const uint32_t full = __ballot_sync(0xFFFFFFFFu, mine_ready == kServerPer);
            const int first_partial = full == 0xFFFFFFFFu ? 32 : __ffs((int)~full) - 1;
            const int partial = first_partial < 32 ? __shfl_sync(0xFFFFFFFFu, mine_ready, first_partial) : 0;
            const int64_t n_ready = min((int64_t)first_partial * kServerPer + partial, p.n_chunks - base);
            if (n_ready == 0) {
                if (p.debug == 2u && ++spins > (1u << 22)) chunk_stuck(p, 5, base, 0);
                continue;
            }
            spins = 0;
            uint64_t v[kServerPer];
            uint64_t mine = 0;
#pragma unroll
            for (int k = 0; k < kServerPer; ++k) {
                const bool take = (int64_t)lane * kServerPer + k < n_ready;
                mine += take ? (d[k] & kValueMask) : 0ull;
                v[k] = mine;
            }
            uint64_t incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint64_t up = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                if (lane >= o) incl += up;
            }
            const uint64_t before = running + incl - mine;
#pragma unroll
            for (int k = 0; k < kServerPer; ++k)
                if ((int64_t)lane * kServerPer + k < n_ready) st_relaxed_gpu(p.status + idx0 + k, kStatusPrefix | (before + v[k]));
            running += __shfl_sync(0xFFFFFFFFu, incl, 31);
            base += n_ready;
        }
        return;
    }

    // ---------------------------------------------------------------------------------------------------- consumers
    const uint32_t lt = lanemask_lt();
    const uint32_t slot_addr0 = smem_u32(slots) + (uint32_t)warp * (128u * 8u) + (uint32_t)lane * 8u;   // this warp's 128 rows of a ring slot
    int slot = 0;
    uint32_t phase = 0;

    // ---- P2 of iteration j: compact the chunk in buffer j % kChunkBufs
    auto compact = [&](int64_t j) {
        const int pb = (int)(j % kChunkBufs);
        const uint32_t ppar = (uint32_t)((j / kChunkBufs) & 1);
        const int64_t pc = sm.chunk_id[pb];
        const bool from_smem = pc * kChunkRows + kChunkRows <= p.n_rows;
        CHUNK_WAIT(mbar_try_wait(&sm.p1done[pb], ppar), 6, j, pc);     // every warp's counts and selection words of that chunk
        uint32_t tc[kChunkTiles];
#pragma unroll
        for (int t = 0; t < kChunkTiles; ++t) {
            tc[t] = 0;
#pragma unroll
            for (int w = 0; w < kChunkCW / kChunkTiles; ++w) tc[t] += sm.wcnt[pb][t * (kChunkCW / kChunkTiles) + w];
        }
        bool have_prefix = false;
        uint64_t prefix = 0;
#pragma unroll 1
        for (int t = 0; t < kChunkTiles; ++t) {
            const int64_t left = tile_rows_left(pc, t);
            if (left <= 0 || tc[t] == 0u) continue;
            const bool whole = left >= kTileRows;
            const bool dense = tc[t] > p.sparse_max && whole;
            const int64_t row0 = pc * kChunkRows + (int64_t)t * kTileRows;
            const int64_t wrow0 = row0 + (int64_t)warp * 128;
            // survivors of the tile in front of this warp's 128 rows (4 words), from the shared selection words
            const uint32_t w0 = sm.sel[pb][t * 64 + lane], w1 = sm.sel[pb][t * 64 + 32 + lane];
            const int first_word = warp * 4;
            const uint32_t c0 = __popc(w0), c1 = __popc(w1);
            uint32_t wfirst;
            if (first_word < 32) wfirst = __reduce_add_sync(0xFFFFFFFFu, lane < first_word ? c0 : 0u);
            else wfirst = __reduce_add_sync(0xFFFFFFFFu, c0) + __reduce_add_sync(0xFFFFFFFFu, lane < first_word - 32 ? c1 : 0u);
            const uint32_t wsel = first_word < 32 ? w0 : w1;
            uint32_t selw[4], off8[4];
            bool kp[4];
            uint32_t wcnt = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                selw[k] = __shfl_sync(0xFFFFFFFFu, wsel, (first_word + k) & 31);
                const uint32_t r = wfirst + wcnt + __popc(selw[k] & lt);
                wcnt += __popc(selw[k]);
                kp[k] = ((selw[k] >> lane) & 1u) != 0u;
                off8[k] = r * 8u;
            }
            for (int col = 0; col < p.n_col8; ++col) {
                const Col8& cc = p.col8[col];
                uint64_t v[4];
                if (col == p.pred_col && from_smem) {
                    const uint64_t* src = &sm.kbuf[pb][t * kTileRows + warp * 128 + lane];
#pragma unroll
                    for (int k = 0; k < 4; ++k) v[k] = src[k * 32];
                } else if (col != p.pred_col && dense && cc.vec_ok) {
                    CHUNK_WAIT(mbar_try_wait(&sm.full[slot], phase), 7, j, slot);
                    const uint32_t src = slot_addr0 + (uint32_t)slot * kSlotBytes;
#pragma unroll
                    for (int k = 0; k < 4; ++k) v[k] = lds64(src + k * 256);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&sm.empty[slot]);
                    if (++slot == p.n_slots) { slot = 0; phase ^= 1u; }
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        v[k] = 0ull;
                        if (kp[k]) v[k] = ld_gather(cc.in + wrow0 + k * 32 + lane);   // rows with a set bit exist
                    }
                }
                if (cc.valid.words != nullptr) {   // placeholder 0 under a null (primitive.rs:175-178)
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t vm = load_bits32(cc.valid, (uint64_t)(wrow0 + k * 32));
                        if (((vm >> lane) & 1u) == 0u) v[k] = 0ull;
                    }
                }
                if (!have_prefix) {
                    CHUNK_WAIT(mbar_try_wait(&sm.pready[pb], ppar), 8, j, pc);
                    prefix = sm.prefix[pb];
                    have_prefix = true;
                }
                uint64_t obase = prefix - base0;
#pragma unroll
                for (int u = 0; u < kChunkTiles; ++u) obase += u < t ? tc[u] : 0u;
                char* ob = reinterpret_cast<char*>(cc.out + obase);
                asm volatile("" : "+l"(ob));
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (kp[k]) st_stream(reinterpret_cast<uint64_t*>(ob + off8[k]), v[k]);
            }
        }
        // Every consumer observes "prefix ready" of the chunk before it lets go of the buffer, survivors or not: the scanner then
        // never falls more than kChunkLag chunks behind, which keeps the parity-tracked barriers (kChunkBufs uses apart) unambiguous.
        if (!have_prefix) CHUNK_WAIT(mbar_try_wait(&sm.pready[pb], ppar), 9, j, pc);
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.kfree[pb]);
    };

#pragma unroll 1
    for (int64_t i = 0;; ++i) {
        const int b = (int)(i % kChunkBufs);
        const uint32_t par = (uint32_t)((i / kChunkBufs) & 1);
        CHUNK_WAIT(mbar_try_wait(&sm.kfull[b], par), 10, i, b);
        const int64_t c = sm.chunk_id[b];
        if (c >= 0) {
            // ---- P1: this warp's 256 rows of chunk c = 8 selection words; lane k keeps word k
            const int64_t wrow0 = c * kChunkRows + (int64_t)warp * 256;
            const bool from_smem = c * kChunkRows + kChunkRows <= p.n_rows;
            uint32_t vwords = 0xFFFFFFFFu;
            if (p.pred_valid.words != nullptr && lane < 8) vwords = load_bits32(p.pred_valid, (uint64_t)(wrow0 + 32 * lane));
            uint32_t myword = 0;
            const uint64_t* const ksrc = &sm.kbuf[b][warp * 256 + lane];
            uint64_t v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (from_smem) v[k] = ksrc[k * 32];
                else { const int64_t row = wrow0 + k * 32 + lane; v[k] = row < p.n_rows ? ld_stream(p.pred_values + row) : 0ull; }
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                bool keep = chunk_keep<PRED>(p, v[k]);
                if (p.pred_valid.words != nullptr) {
                    const uint32_t vw = __shfl_sync(0xFFFFFFFFu, vwords, k);
                    keep = ((vw >> lane) & 1u) ? keep : (p.keep_null != 0u);
                }
                if (!from_smem && wrow0 + k * 32 + lane >= p.n_rows) keep = false;
                const uint32_t m = __ballot_sync(0xFFFFFFFFu, keep);
                if (lane == k) myword = m;
            }
            if (lane < 8) {
                sm.sel[b][warp * 8 + lane] = myword;
                // whole tiles of the selection bitmap exist (zero beyond the last row): the kernels that follow read 64 words per tile
                if (tile_rows_left(c, warp / (kChunkCW / kChunkTiles)) > 0) p.sel_out[((c * kChunkRows) >> 5) + warp * 8 + lane] = myword;
            }
            const uint32_t wc = __reduce_add_sync(0xFFFFFFFFu, lane < 8 ? (uint32_t)__popc(myword) : 0u);
            __syncwarp();
            if (lane == 0) {
                sm.wcnt[b][warp] = wc;
                atomicAdd(&sm.ctot[b], wc);
                __threadfence_block();
                if (atomicAdd(&sm.arrived[b], 1u) == kChunkCW - 1) {
                    // last consumer warp of this chunk's P1: every count is in — publish the chunk's descriptor for the prefix server
                    __threadfence_block();
                    const uint32_t total = *reinterpret_cast<volatile uint32_t*>(&sm.ctot[b]);
                    st_relaxed_gpu(p.status + c, kStatusAggregate | (uint64_t)total);
                    sm.ctot[b] = 0u; sm.arrived[b] = 0u;   // the buffer's next chunk is kChunkBufs iterations away
                }
                mbar_arrive(&sm.p1done[b]);
            }
        }
        if (i >= kChunkLag) compact(i - kChunkLag);
        if (c < 0) {
            // no chunk left: drain the ones still waiting for their P2
            for (int64_t j = (i >= kChunkLag ? i - kChunkLag + 1 : 0); j < i; ++j) compact(j);
            break;
        }
    }
}

// Selectivity estimate for the plan choice: `groups` runs of 64 consecutive rows spread evenly over the batch (64 K rows by default,
// ~5 us), the predicate evaluated exactly as the plans do.  The single-pass chunk plan wins from ~30 % survivors upwards (it saves the
// second read of the predicate column, worth the more the denser the tiles are) and loses below (its pipeline is latency-bound when
// there is little to compact), so a blocking call spends one tiny kernel + count readback on knowing which side it is on.
template <int PRED>
__global__ void __launch_bounds__(kBlock) sample_selectivity_kernel(const __grid_constant__ ChunkParams p, int64_t groups, unsigned long long* out) {
    const int lane = threadIdx.x & 31;
    const int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (g >= groups) return;
    const int64_t start = ((g * (p.n_rows / groups)) >> 6) << 6;
    uint32_t kept = 0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int64_t row = start + h * 32 + lane;
        bool keep = false;
        if (row < p.n_rows) {
            keep = chunk_keep<PRED>(p, ld_stream(p.pred_values + row));
            if (p.pred_valid.words != nullptr) {
                const uint64_t bit = p.pred_valid.bit0 + (uint64_t)row;
                if (((__ldg(p.pred_valid.words + (bit >> 5)) >> (bit & 31)) & 1u) == 0u) keep = p.keep_null != 0u;
            }
        }
        kept += __popc(__ballot_sync(0xFFFFFFFFu, keep));
    }
    if (lane == 0 && kept != 0u) atomicAdd(out, (unsigned long long)kept);
}

}  // namespace rvl
