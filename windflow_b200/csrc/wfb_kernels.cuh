// wfb_kernels.cuh -- hand-written sm_100a kernels of the WindFlow GPU operator hot path, templated on a
// "program" (record schema + functors, see wfb_programs.cuh).
//
// Kernel inventory (DESIGN.md sections 3-4 have the algorithms and the measured roofline of each):
//   k_tile_pass<P, MODE>     streaming tile pass, persistent warp-specialised CTAs: TMA load of a tile of tuples
//                            (cp.async.bulk.tensor / cp.async.bulk + mbarrier rings), per-tuple map / filter / lift / key->slot
//                            in registers, survivors staged in shared memory, TMA bulk store.
//                              MODE_MAP     in-place Map_GPU                 (wf/map_gpu.hpp:61-76)
//                              MODE_FILTER  [Map_GPU ->] Filter_GPU, compacted per batch with a decoupled look-back
//                                           (wf/filter_gpu.hpp:72-88, :555-570)
//                              MODE_INGEST  [Map -> Filter ->] lift + key->slot for Ffat_Windows_GPU, or -> destination for
//                                           wfb_shard_lift (wf/ffat_replica_gpu.hpp:94-121); chain-free ("sparse") on the bucket path
//   k_wide_tile_hist / k_wide_scatter / k_shard_scatter   one stable partition pass on a 10-bit digit (window path: 1024
//                            buckets of slots; multi-GPU: destinations, with the records)
//   k_ffat_update_buckets<P> one CTA per bucket: split by key, ordered pane folds, FlatFAT leaf + path update, fired
//                            groups (wf/flatfat_gpu.hpp:62-139, wf/ffat_replica_gpu.hpp:830-867); k_ffat_windows<P> window queries
//   k_onesweep_pass / k_radix_ghist   LSD radix sort, 8 bits per pass, one kernel per pass: the replacement of
//                            thrust::sort_by_key for the per-batch keyed operators and the full-sort window path
//                            (wf/ffat_replica_gpu.hpp:751, wf/keyby_emitter_gpu.hpp:547, wf/reduce_gpu.hpp:239)
//   k_ffat_update_lanes / k_ffat_update   full-sort window path: one thread / one warp per key
//   k_extract_keys, k_seg_*, k_reduce_segments*, k_reduce_all, k_gather_tuples   KeyBy_Emitter_GPU grouping, Reduce_GPU
#pragma once
#include <type_traits>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include "wfb_ptx.cuh"

namespace wfb {

constexpr int TILE = 256;    // tuples per tile == threads per CTA of k_tile_pass
constexpr uint32_t OSW_TILE_POS = 4096; // positions per tile of the wide partition pass (== OSW_TILE below)
#ifndef WFB_STAGES
#define WFB_STAGES 4
#endif
constexpr int STAGES = WFB_STAGES;    // TMA ring depth per CTA
constexpr uint32_t FULL = 0xffffffffu;

enum { MODE_MAP = 0, MODE_FILTER = 1, MODE_INGEST = 2 };
constexpr uint32_t MAX_SHARDS = 8;

// decoupled look-back tile state: [63:34] epoch, [33:32] status, [31:0] value
constexpr uint64_t ST_AGG = 1, ST_PREFIX = 2;
__host__ __device__ __forceinline__ uint64_t pack_state(uint32_t epoch, uint64_t status, uint32_t v)
{
    return (static_cast<uint64_t>(epoch & 0x3fffffffu) << 34) | (status << 32) | v;
}

// one input batch as the kernels see it
struct DevBatch {
    const unsigned char *tuples;
    const uint64_t *ts;
    unsigned char *out;        // MODE_FILTER/MAP: output tuples
    uint64_t *ts_out;          // MODE_FILTER: output timestamps (may be null)
    uint32_t *n_out;           // MODE_FILTER: survivors of this batch (device)
    uint64_t watermark;
    uint32_t n;
    uint32_t tile_begin;       // first global tile index of this batch
};

// one fired window group whose Nb window queries are evaluated after the update kernel
struct Trigger { uint64_t key; uint64_t g; uint32_t slot; uint32_t last_pos; uint32_t obase; uint32_t pad; };

// key -> slot table (open addressing, linear probing) + per-key window state of one Ffat_Windows_GPU
struct FfatDev {
    // key table
    uint64_t *ht_keys;         // capacity entries, EMPTY_KEY when free
    uint32_t *ht_slots;        // capacity entries, INVALID_SLOT until published
    uint32_t ht_mask;          // capacity - 1
    uint32_t max_keys;
    uint32_t *n_slots;         // number of keys inserted so far
    uint32_t *err_flags;       // bit0: key table full, bit1: output capacity exceeded
    unsigned long long *results_total; // window results delivered so far (added by the last kernel of every call)
    uint64_t defer_items;              // a fired group may wait for the deferred pass (k_ffat_windows*) while fewer than this many further items of
                                       // its key follow in the call: (spare ring leaves + 1) panes -- later panes then do not overwrite leaves it reads
    uint32_t dense;            // 1: slot = key (keys < max_keys), or key / key_div for one shard of a keyby
    uint32_t key_div, key_rem; // dense: the handle owns the keys with key % key_div == key_rem (key_div <= 1: all keys)
    // per-slot state
    uint64_t *slot_key;        // key of each slot
    uint64_t *cnt;             // lifted results appended so far (Key_Descriptor::count)
    unsigned char *acc;        // open-pane accumulator, result_t per slot
    unsigned char *tree;       // FlatFAT per slot: (2*n_leaves-1) result_t, leaves first (level 0), root last
    uint32_t *seg_cnt;         // full-sort path only: items of the current stream segment per slot (TileArgs::count_keys; zeroed by the
                               // update kernels). The bucket path counts per key inside the bucket CTA instead.
    uint32_t *seg_off;         // exclusive offsets into the sorted segment, max_keys+1
    struct Trigger *trig;      // deferred window groups of the current segment (evaluated by k_ffat_windows)
    uint32_t *n_trig;          // number of deferred groups
    uint32_t trig_cap;
    uint32_t *heavy;           // slots with more than light_max items in the segment (handled warp-per-key)
    uint32_t *n_heavy;
    uint32_t light_max;
    // window geometry (in tuples and in panes)
    uint64_t win, slide, B;    // B = (Nb-1)*slide + win  (ffat_replica_gpu.hpp:657)
    uint32_t nb;               // windows per trigger
    uint32_t pane;             // pane length P = gcd(win, slide) in tuples
    uint32_t wp, sp;           // window / slide in panes
    uint32_t n_leaves;         // power of two >= B / P
    uint32_t log_leaves;
    uint32_t lazy;             // 1: the update kernels only write the pane LEAVES of the key's FlatFAT; the internal levels a group of windows needs
                               // are built in shared memory when the group is evaluated (k_ffat_windows_lazy). A pane completes once per
                               // `pane` items but fires windows only once per slide * Nb items: maintaining log2(n) path nodes per pane in
                               // global memory costs more than rebuilding n - 1 nodes on chip per fired group.
};
constexpr uint64_t EMPTY_KEY = 0xffffffffffffffffull;
constexpr uint32_t INVALID_SLOT = 0xffffffffu;

struct TileArgs {
    const DevBatch *batches;   // device array (nbatches entries) or null => use `one`
    DevBatch one;
    uint32_t nbatches;
    uint32_t num_tiles;
    uint64_t *tile_state;      // num_tiles words (epoch-tagged, never cleared)
    uint32_t *ticket;          // monotonically increasing ticket counter
    uint32_t ticket_base;      // value of *ticket when this launch starts
    uint32_t epoch;
    uint64_t tmap_base;        // global address the 2-D tensor map starts at (rows of 64 bytes)
    uint32_t use_tmap;         // 1: `tmap` is valid for this launch
    uint32_t max_ctas_per_sm;  // host-side launch hint (0 = no limit), not read by the kernel
    uint32_t sparse;           // MODE_INGEST, 1: no global compaction -- tile t owns lifted / slots [t*TILE, +TILE) (survivors first,
                               // INVALID_SLOT padding), so tiles are independent: no look-back chain, positions are tuple indices
    uint32_t count_keys;       // MODE_INGEST: 1 = per-key item counts of the segment (ff.seg_cnt) for the full-sort update kernels
    uint32_t *wide_h32;        // MODE_INGEST + sparse: per-tile digit counts of the wide partition that follows ([position / 4096][1024], 32-bit),
                               // accumulated here so that the partition needs no counting pass of its own (null: it counts itself)
    uint16_t *wide_h16;        // MODE_INGEST + sparse: per-tile digit counts of the wide partition ([position / 4096][1024], 16-bit rows), filed by
                               // the tile pass itself: with tiles_per_ticket = 16 a CTA owns whole wide tiles, counts their digits in shared
                               // memory and writes each row once -- the partition then needs no counting pass (k_wide_tile_hist) of its own
    uint32_t tiles_per_ticket; // consecutive tiles a producer claims per ticket (0 or 1: one; 16 with wide_h16)
    uint32_t pack_rank;        // with wide_h16 and at most 65536 slots: slots[pos] = slot | rank << 16, rank = what the digit counter of the wide
                               // tile held when this survivor was counted (any order among the survivors of one digit): k_wide_scatter_ranked
                               // places the pairs with it and then restores arrival order inside every (wide tile, digit) cell
    const uint32_t *ext_slots; // MODE_INGEST + in-place: slot of the record at every position, given by the caller (the time-based
                               // front end knows the key slot of every pane it pops); the program's key extractor is not used
    uint32_t inplace;          // MODE_INGEST + sparse, 1: the program passes records through unchanged (lift = identity, no map) and the
                               // batches lie at their tile positions in one buffer: nothing is copied, only the slots are written
    uint32_t l2_hints;         // 1: input tiles are loaded evict-first, lifted records stored evict-last (they are re-read by the update)
    // MODE_INGEST with nshards != 0 (wfb_shard_lift): the "slot" of a tuple is its destination key % nshards
    uint32_t nshards, region_cap;
    // ... and with shard_slots != 0 (bucketed exchange, wfb_mg_*): the "slot" is the VIRTUAL slot dest * shard_slots + key / nshards
    // (shard_slots a power of two, nshards * shard_slots <= 65536), so that ONE wide partition at the source leaves the records
    // grouped by (destination, bucket of the destination's slot space); keys at or above shard_keys * nshards raise *shard_err
    uint32_t shard_slots, shard_keys;
    uint32_t *shard_err;
    // MODE_INGEST: digit histograms of the slot sort that follows (RadixSorter ctl), accumulated per CTA in shared memory
    uint32_t *sort_ctl; uint32_t sort_passes, sort_shift, sort_dbits; // sort_dbits: digit width of a pass (8, or 10 for the wide pass)
    // MODE_INGEST outputs (compacted over the whole segment, arrival order)
    unsigned char *lifted;     // result_t per surviving tuple
    uint32_t *slots;           // slot per surviving tuple
    uint32_t *batch_off;       // nbatches+1: compact offset of the first survivor of each batch; [nbatches]=total
    uint32_t *n_total;         // == batch_off[nbatches]
    FfatDev ff;
};

// ------------------------------------------------------------------------------------------------------
// shared-memory tile access. A tuple of C = sizeof(T)/16 16-byte chunks is read/written with its chunks
// rotated by (idx*C/8)%C so that the 8 lanes of a quarter warp hit 8 different 16-byte bank groups
// (stride-64B LDS.128 would otherwise be a 4-way bank conflict).
// ------------------------------------------------------------------------------------------------------
template <class T>
struct TileIO {
    static constexpr int TB = sizeof(T);
    static constexpr bool V16 = (TB % 16 == 0) && (TB / 16 == 1 || TB / 16 == 2 || TB / 16 == 4 || TB / 16 == 8);
    static constexpr int C = V16 ? TB / 16 : TB / 8;

    __device__ __forceinline__ static void load(const unsigned char *base, uint32_t idx, T &out)
    {
        if constexpr (V16) {
            uint4 v[C];
            const uint4 *p = reinterpret_cast<const uint4 *>(base + static_cast<size_t>(idx) * TB);
            const uint32_t rot = (C > 1) ? ((idx * C / 8) % C) : 0;
#pragma unroll
            for (int j = 0; j < C; j++) v[j] = p[(j + rot) & (C - 1)];
            // v[j] holds chunk (j+rot)%C; rotate right by rot so that v[k] holds chunk k
#pragma unroll
            for (int s = 1; s < C; s <<= 1) {
                if (rot & s) {
                    uint4 t[C];
#pragma unroll
                    for (int k = 0; k < C; k++) t[k] = v[(k - s) & (C - 1)];
#pragma unroll
                    for (int k = 0; k < C; k++) v[k] = t[k];
                }
            }
            uint4 *o = reinterpret_cast<uint4 *>(&out);
#pragma unroll
            for (int k = 0; k < C; k++) o[k] = v[k];
        } else {
            const uint64_t *p = reinterpret_cast<const uint64_t *>(base + static_cast<size_t>(idx) * TB);
            uint64_t *o = reinterpret_cast<uint64_t *>(&out);
#pragma unroll
            for (int k = 0; k < C; k++) o[k] = p[k];
        }
    }

    __device__ __forceinline__ static void store(unsigned char *base, uint32_t idx, const T &in)
    {
        if constexpr (V16) {
            uint4 v[C];
            const uint4 *src = reinterpret_cast<const uint4 *>(&in);
#pragma unroll
            for (int k = 0; k < C; k++) v[k] = src[k];
            const uint32_t rot = (C > 1) ? ((idx * C / 8) % C) : 0;
            // rotate left by rot: v'[j] = chunk (j+rot)%C, then store v'[j] at position (j+rot)%C
#pragma unroll
            for (int s = 1; s < C; s <<= 1) {
                if (rot & s) {
                    uint4 t[C];
#pragma unroll
                    for (int k = 0; k < C; k++) t[k] = v[(k + s) & (C - 1)];
#pragma unroll
                    for (int k = 0; k < C; k++) v[k] = t[k];
                }
            }
            uint4 *p = reinterpret_cast<uint4 *>(base + static_cast<size_t>(idx) * TB);
#pragma unroll
            for (int j = 0; j < C; j++) p[(j + rot) & (C - 1)] = v[j];
        } else {
            uint64_t *p = reinterpret_cast<uint64_t *>(base + static_cast<size_t>(idx) * TB);
            const uint64_t *s = reinterpret_cast<const uint64_t *>(&in);
#pragma unroll
            for (int k = 0; k < C; k++) p[k] = s[k];
        }
    }
};

// global <-> shared tile movement: TMA bulk copy when size and address allow it, else coalesced 8-byte words
__device__ __forceinline__ bool bulk_ok(const void *g, uint32_t bytes)
{
    return ((reinterpret_cast<uintptr_t>(g) | bytes) & 15u) == 0;
}

// ------------------------------------------------------------------------------------------------------
// key -> slot lookup / insert (replaces the host unordered_map of ffat_replica_gpu.hpp:514, :783-795)
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t mix64(uint64_t x)
{
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
}

__device__ __forceinline__ uint64_t key_of_slot(const FfatDev &ff, uint32_t slot)
{
    return ff.dense ? (ff.key_div > 1 ? static_cast<uint64_t>(slot) * ff.key_div + ff.key_rem : static_cast<uint64_t>(slot)) : ff.slot_key[slot];
}

__device__ __forceinline__ uint32_t slot_of_key(const FfatDev &ff, uint64_t key)
{
    if (ff.dense) {
        if (ff.key_div > 1) { // one shard of a keyby: keys with key % key_div == key_rem, compact slots
            if (key % ff.key_div != ff.key_rem) { atomicOr(ff.err_flags, 1u); return INVALID_SLOT; }
            key /= ff.key_div;
        }
        if (key >= ff.max_keys) { atomicOr(ff.err_flags, 1u); return INVALID_SLOT; }
        return static_cast<uint32_t>(key);
    }
    uint32_t h = static_cast<uint32_t>(mix64(key)) & ff.ht_mask;
    for (uint32_t probe = 0; probe <= ff.ht_mask; probe++) {
        uint64_t k = ld_relaxed_u64(&ff.ht_keys[h]);
        if (k == EMPTY_KEY) {
            unsigned long long old = atomicCAS(reinterpret_cast<unsigned long long *>(&ff.ht_keys[h]),
                                               static_cast<unsigned long long>(EMPTY_KEY),
                                               static_cast<unsigned long long>(key));
            if (old == EMPTY_KEY) { // we own the entry: allocate the slot and publish it
                uint32_t s = atomicAdd(ff.n_slots, 1u);
                if (s >= ff.max_keys) { atomicOr(ff.err_flags, 1u); s = INVALID_SLOT - 1; }
                else ff.slot_key[s] = key;
                st_release_u32(&ff.ht_slots[h], s);
                return s >= ff.max_keys ? INVALID_SLOT : s;
            }
            k = old;
        }
        if (k == key) {
            uint32_t s;
            while ((s = ld_acquire_u32(&ff.ht_slots[h])) == INVALID_SLOT) { }
            return s >= ff.max_keys ? INVALID_SLOT : s;
        }
        h = (h + 1) & ff.ht_mask;
    }
    atomicOr(ff.err_flags, 1u);
    return INVALID_SLOT;
}

// ------------------------------------------------------------------------------------------------------
// k_tile_pass: persistent, warp-specialised CTAs (10 warps), dynamic tile tickets, STAGES-deep TMA ring.
//   warp 0      PRODUCER  (one lane): wait empty[s] | claim ticket | find the batch | TMA load of the tile:
//                         cp.async.bulk.tensor.2d with SWIZZLE_64B for full tiles of 64-byte tuples (bank-conflict
//                         free LDS.128 without register shuffling), cp.async.bulk (linear) otherwise; timestamps of
//                         the tile ride on the same mbarrier.
//   warps 1..8  CONSUMERS (one tuple per thread): wait full[s] | tuple -> registers | map | filter | [lift,
//                         key->slot, per-key count] | ballot + warp totals -> local offsets (one named barrier) |
//                         survivors -> shared (compacted, linear) | publish the tile count (look-back AGGREGATE) |
//                         arrive staged[s] and move on to the next tile.
//   warp 9      EPILOGUE: wait staged[s] | decoupled look-back -> tile base, publish PREFIX | one TMA bulk store of
//                         the compacted records, coalesced stores of the staged slots / timestamps | wait for the
//                         store to have read shared memory | arrive empty[s].
// Tickets: every CTA claims one ticket per processed tile plus the failing one, so a launch consumes exactly
// num_tiles + gridDim.x tickets (the host advances ticket_base by that amount). (Claiming several tiles per atomic was
// measured: it delays the aggregates the look-backs of the following tiles wait for, 0.10 -> 0.14 ms.)
// ------------------------------------------------------------------------------------------------------
constexpr uint32_t TP_THREADS = TILE + 64;         // producer warp + 8 consumer warps + epilogue warp
constexpr uint32_t TILE_SENTINEL = 0x7fffffffu;
enum { TF_SWZ = 1u, TF_FALLBACK = 2u, TF_TS_SMEM = 4u, TF_WIDE_LAST = 8u };

struct StageMeta {
    uint32_t tile, batch, first, cnt, flags, count; // count: survivors (written by the consumers)
    uint32_t wide, pad;                             // wide: ticket (= wide tile when tiles_per_ticket = 16)
};

template <class P, int MODE>
struct TilePassSmem {
    using T = typename P::tuple_t;
    using R = typename P::result_t;
    static constexpr uint32_t rec_bytes = (MODE == MODE_INGEST && sizeof(R) > sizeof(T)) ? sizeof(R) : sizeof(T);
    static constexpr uint32_t tile_bytes = (TILE * rec_bytes + 1023u) & ~1023u; // swizzled stages need 512-B alignment
    static constexpr uint32_t aux_bytes = (MODE == MODE_INGEST) ? TILE * 4u : (MODE == MODE_FILTER ? TILE * 16u : 0u); // slots | ts in + ts out
    static constexpr uint32_t stage_bytes = tile_bytes + aux_bytes;
    static constexpr uint32_t hist_bytes = (MODE == MODE_INGEST) ? 4u * 256u * 4u : 0u; // up to 4 sort passes of 8 bits or one of 10 | two buffers of 1024 packed 16-bit per-wide-tile counts
    static constexpr uint32_t total = STAGES * stage_bytes + 1024 /*alignment slack*/ + 1024 /*barriers, meta, scan*/ + hist_bytes;
};

__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const void *tmap, int32_t c0, int32_t c1, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_2d_hint(void *smem_dst, const void *tmap, int32_t c0, int32_t c1, uint64_t *bar, uint64_t policy)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
                 ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(policy) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const void *tmap, int32_t c0, int32_t c1, const void *smem_src)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(tmap), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void consumer_bar() { asm volatile("bar.sync 1, %0;" ::"n"(TILE) : "memory"); }

template <class P, int MODE>
__global__ void __launch_bounds__(TP_THREADS) k_tile_pass(const __grid_constant__ CUtensorMap tmap, const TileArgs a,
                                                          const __grid_constant__ typename P::params_t prm)
{
    using T = typename P::tuple_t;
    using R = typename P::result_t;
    using SM = TilePassSmem<P, MODE>;
    constexpr uint32_t TB = sizeof(T);
    constexpr uint32_t RB = sizeof(R);
    static_assert(TB % 8 == 0 && RB % 8 == 0, "records must be multiples of 8 bytes");
    constexpr bool CAN_SWZ = (TB == 64);

    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    unsigned char *ctl = smem + STAGES * SM::stage_bytes;
    uint64_t *full = reinterpret_cast<uint64_t *>(ctl);          // STAGES  producer -> consumers (tx)
    uint64_t *staged = full + STAGES;                            // STAGES  consumers -> epilogue
    uint64_t *empty = staged + STAGES;                           // STAGES  epilogue -> producer
    StageMeta *meta = reinterpret_cast<StageMeta *>(empty + STAGES);  // STAGES
    uint32_t *warp_tot = reinterpret_cast<uint32_t *>(meta + STAGES); // 2 x 8 warp totals (double-buffered by iteration parity)
    uint32_t *s_hist = reinterpret_cast<uint32_t *>(ctl + 1024);      // MODE_INGEST: [pass][1 << sort_dbits] digit counts of this CTA (1024 words)

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    auto stage_buf = [&](uint32_t s) { return smem + s * SM::stage_bytes; };
    auto stage_aux = [&](uint32_t s) { return smem + s * SM::stage_bytes + SM::tile_bytes; };

    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&staged[s], TILE / 32); mbar_init(&empty[s], 1); }
        mbar_fence_init();
    }
    if constexpr (MODE == MODE_INGEST) { for (uint32_t i = tid; i < 4u * 256u; i += TP_THREADS) s_hist[i] = 0; }
    __syncthreads();

    if (warp == 0) {
        // ================================= PRODUCER =================================
        if (lane == 0) {
            const uint64_t pol_first = l2_policy_evict_first();
            DevBatch b = a.one; uint32_t bi = 0, b_end = 0; // batch of the previous tile, first tile after it
            bool have_batch = false;
            const uint32_t tpt = a.tiles_per_ticket ? a.tiles_per_ticket : 1u;
            uint32_t t_next = 0, t_end = 0, claim = 0; // tiles [t_next, t_end) of the current claim are still to load
            for (uint32_t it = 0;; it++) {
                const uint32_t s = it % STAGES, par = (it / STAGES) & 1u;
                mbar_wait(&empty[s], par ^ 1u); // a fresh barrier passes the wait on parity 1
                // claim only now: a claimed tile is loaded at once, so the look-backs of its successors never wait on a stalled ring
                StageMeta &m = meta[s];
                if (t_next == t_end) {
                    claim = atomicAdd(a.ticket, 1u) - a.ticket_base;
                    t_next = claim * tpt; // (claim <= (num_tiles + grid) / tpt: no overflow)
                    if (claim >= (a.num_tiles + tpt - 1) / tpt) { m.tile = TILE_SENTINEL; mbar_arrive(&full[s]); break; }
                    t_end = min(t_next + tpt, a.num_tiles);
                }
                const uint32_t t = t_next++;
                if (a.batches != nullptr && !(have_batch && t >= b.tile_begin && t < b_end)) {
                    uint32_t lo = 0, hi = a.nbatches - 1; // last batch with tile_begin <= t
                    while (lo < hi) { const uint32_t mid = (lo + hi + 1) >> 1; if (a.batches[mid].tile_begin <= t) lo = mid; else hi = mid - 1; }
                    bi = lo; b = a.batches[lo];
                    b_end = (lo + 1 < a.nbatches) ? a.batches[lo + 1].tile_begin : a.num_tiles;
                    have_batch = true;
                }
                const uint32_t first = (t - b.tile_begin) * TILE;
                const uint32_t cnt = min(static_cast<uint32_t>(TILE), b.n - first);
                const unsigned char *src = b.tuples + static_cast<size_t>(first) * TB;
                uint32_t flags = 0, tx = 0;
                const uint64_t off = reinterpret_cast<uint64_t>(src) - a.tmap_base;
                if (CAN_SWZ && a.use_tmap && cnt == TILE && (off & 63u) == 0 && (off >> 6) < 0x7fffff00ull) flags = TF_SWZ;
                else if (!bulk_ok(src, cnt * TB)) flags = TF_FALLBACK;
                if (!(flags & TF_FALLBACK)) tx += cnt * TB;
                const uint64_t *tsp = (MODE == MODE_FILTER && b.ts != nullptr) ? b.ts + first : nullptr;
                if (tsp != nullptr && bulk_ok(tsp, cnt * 8u)) { flags |= TF_TS_SMEM; tx += cnt * 8u; }
                if (t_next == t_end) flags |= TF_WIDE_LAST;
                m.tile = t; m.batch = bi; m.first = first; m.cnt = cnt; m.flags = flags; m.wide = claim;
                if (tx) mbar_expect_tx(&full[s], tx); else mbar_arrive(&full[s]);
                if (a.l2_hints) { // read-once stream: first in line for eviction
                    if (flags & TF_SWZ) tma_load_2d_hint(stage_buf(s), &tmap, 0, static_cast<int32_t>(off >> 6), &full[s], pol_first);
                    else if (!(flags & TF_FALLBACK)) bulk_g2s_hint(stage_buf(s), src, cnt * TB, &full[s], pol_first);
                } else if (flags & TF_SWZ) tma_load_2d(stage_buf(s), &tmap, 0, static_cast<int32_t>(off >> 6), &full[s]);
                else if (!(flags & TF_FALLBACK)) bulk_g2s(stage_buf(s), src, cnt * TB, &full[s]);
                if (flags & TF_TS_SMEM) bulk_g2s(stage_aux(s), tsp, cnt * 8u, &full[s]);
            }
        }
    } else if (warp <= TILE / 32) {
        // ================================= CONSUMERS =================================
        const uint32_t ctid = tid - 32, cwarp = warp - 1;
        uint32_t wpar = 0; // which of the two digit-count buffers the current wide tile uses
        for (uint32_t it = 0;; it++) {
            const uint32_t s = it % STAGES, par = (it / STAGES) & 1u;
            unsigned char *buf = stage_buf(s);
            mbar_wait(&full[s], par);
            const StageMeta m = meta[s];
            if (m.tile == TILE_SENTINEL) { // pass the end-of-work marker on to the epilogue warp
                __syncwarp();
                if (lane == 0) mbar_arrive(&staged[s]);
                if constexpr (MODE == MODE_INGEST) {
                    if (a.sort_ctl != nullptr) { // every consumer is done counting: add this CTA's digit counts to the global ones
                        consumer_bar();
                        for (uint32_t i = ctid; i < (a.sort_passes << a.sort_dbits); i += TILE) { const uint32_t c = s_hist[i]; if (c) atomicAdd(&a.sort_ctl[i], c); }
                    }
                }
                break;
            }
            DevBatch b;
            if (a.batches == nullptr) b = a.one; else b = a.batches[m.batch];
            const uint32_t cnt = m.cnt;
            const bool active = ctid < cnt;
            if (m.flags & TF_FALLBACK) { // unaligned / odd-sized tile: coalesced 8-byte copies into the linear buffer
                const uint64_t *src = reinterpret_cast<const uint64_t *>(b.tuples + static_cast<size_t>(m.first) * TB);
                uint64_t *dst = reinterpret_cast<uint64_t *>(buf);
                for (uint32_t w = ctid; w < cnt * (TB / 8); w += TILE) dst[w] = src[w];
                consumer_bar();
            }
            // ---- per-tuple work in registers ---------------------------------------------------------------
            alignas(16) T tup;
            uint64_t ts = 0;
            bool keep = false;
            if (active) {
                if constexpr (CAN_SWZ) {
                    if (m.flags & TF_SWZ) { // SWIZZLE_64B: 16-byte chunk j of row r lives at chunk j ^ ((r >> 1) & 3)
                        const uint4 *p = reinterpret_cast<const uint4 *>(buf + static_cast<size_t>(ctid) * 64);
                        const uint32_t x = (ctid >> 1) & 3u;
                        uint4 *o = reinterpret_cast<uint4 *>(&tup);
#pragma unroll
                        for (uint32_t jj = 0; jj < 4; jj++) o[jj] = p[jj ^ x];
                    } else TileIO<T>::load(buf, ctid, tup);
                } else TileIO<T>::load(buf, ctid, tup);
                if constexpr (MODE == MODE_FILTER) {
                    if (b.ts != nullptr)
                        ts = (m.flags & TF_TS_SMEM) ? reinterpret_cast<const uint64_t *>(stage_aux(s))[ctid] : b.ts[m.first + ctid];
                }
                P::map(tup, prm);
                keep = (MODE == MODE_MAP) ? true : P::filter(tup, prm);
            }
            uint32_t slot = INVALID_SLOT;
            alignas(16) R res;
            if constexpr (MODE == MODE_INGEST) {
                if (keep) {
                    P::lift(tup, res, prm);
                    if (a.nshards && a.shard_slots) { // keyby across GPUs, bucketed: destination-major virtual slot
                        const uint64_t key = P::key(tup, prm), q = key / a.nshards;
                        if (q < a.shard_keys) slot = static_cast<uint32_t>(key - q * a.nshards) * a.shard_slots + static_cast<uint32_t>(q);
                        else atomicOr(a.shard_err, 2u); // (a key outside the declared key space: the record is dropped, the step fails)
                    } else if (a.nshards) slot = static_cast<uint32_t>(P::key(tup, prm) % a.nshards); // keyby across GPUs: the "slot" is the destination
                    else {
                        if (a.ext_slots != nullptr) { slot = a.ext_slots[m.tile * TILE + ctid]; if (slot >= a.ff.max_keys) slot = INVALID_SLOT; }
                        else slot = slot_of_key(a.ff, P::key(tup, prm));
                        if (slot != INVALID_SLOT && a.count_keys) atomicAdd(&a.ff.seg_cnt[slot], 1u); // (full-sort path only)
                    }
                    if (a.wide_h16 != nullptr) { // digit counts of this wide tile (the CTA owns all of its tiles)
                        if (slot != INVALID_SLOT) { // two 16-bit counters per word (a wide tile holds 4096 positions: no carry)
                            const uint32_t d = (slot >> a.sort_shift) & 1023u;
                            const uint32_t before = atomicAdd(&s_hist[wpar * 512u + (d >> 1)], 1u << ((d & 1u) * 16u));
                            if (a.pack_rank) slot |= ((before >> ((d & 1u) * 16u)) & 0xffffu) << 16; // (slot < 65536: the rank rides in the upper half)
                        }
                    } else if (a.sort_ctl != nullptr && !(a.sparse && slot == INVALID_SLOT)) { // digit counts for the radix passes over the slots (invalid slots sort last / are skipped)
                        for (uint32_t ps = 0; ps < a.sort_passes; ps++)
                            atomicAdd(&s_hist[(ps << a.sort_dbits) + ((slot >> (a.sort_shift + a.sort_dbits * ps)) & ((1u << a.sort_dbits) - 1u))], 1u);
                        if (a.wide_h32 != nullptr) // (tile t owns positions [256 t, +256): wide tile t / 16)
                            atomicAdd(&a.wide_h32[static_cast<size_t>(m.tile / (OSW_TILE_POS / TILE)) * 1024u + ((slot >> a.sort_shift) & 1023u)], 1u);
                    }
                }
            }
            if constexpr (MODE == MODE_MAP) {
                consumer_bar(); // every consumer has read its tuple: overwrite the stage with the results
                if (active) {
                    if constexpr (CAN_SWZ) {
                        if (m.flags & TF_SWZ) {
                            uint4 *p = reinterpret_cast<uint4 *>(buf + static_cast<size_t>(ctid) * 64);
                            const uint32_t x = (ctid >> 1) & 3u;
                            const uint4 *o = reinterpret_cast<const uint4 *>(&tup);
#pragma unroll
                            for (uint32_t jj = 0; jj < 4; jj++) p[jj ^ x] = o[jj];
                        } else TileIO<T>::store(buf, ctid, tup);
                    } else TileIO<T>::store(buf, ctid, tup);
                }
                if (ctid == 0) meta[s].count = cnt;
            } else {
                // ---- stable local offsets: ballot + warp totals (double-buffered), ONE named barrier ------------
                const uint32_t bal = __ballot_sync(FULL, keep);
                uint32_t *wt = warp_tot + (it & 1u) * (TILE / 32);
                if (lane == 0) wt[cwarp] = __popc(bal);
                consumer_bar(); // totals visible; every consumer has read its tuple (stage re-usable for staging)
                if constexpr (MODE == MODE_INGEST) {
                    if (a.wide_h16 != nullptr && (m.flags & TF_WIDE_LAST)) { // last tile of the wide tile: file its row, clear the buffer for the
                        uint32_t *hrow = s_hist + wpar * 512u;                 // wide tile after the next (a barrier per tile lies in between)
                        const uint2 c = reinterpret_cast<const uint2 *>(hrow)[ctid];
                        reinterpret_cast<uint2 *>(hrow)[ctid] = make_uint2(0, 0);
                        reinterpret_cast<uint2 *>(a.wide_h16 + static_cast<size_t>(m.wide) * 1024u)[ctid] = c; // (little endian: the packed words are the row)
                        wpar ^= 1u;
                    }
                }
                uint32_t wbase = 0, total = 0;
#pragma unroll
                for (uint32_t w = 0; w < TILE / 32; w++) { const uint32_t c = wt[w]; if (w < cwarp) wbase += c; total += c; }
                const uint32_t local = wbase + __popc(bal & lanemask_lt());
                if (ctid == 0) { // publish the aggregate right away so that other CTAs' look-backs never wait for us
                    meta[s].count = total;
                    const uint32_t chain_begin = (MODE == MODE_FILTER) ? b.tile_begin : 0u;
                    if (m.tile != chain_begin && !(MODE == MODE_INGEST && a.sparse)) st_relaxed_u64(&a.tile_state[m.tile], pack_state(a.epoch, ST_AGG, total));
                }
                if (keep) {
                    if constexpr (MODE == MODE_FILTER) {
                        TileIO<T>::store(buf, local, tup);
                        if (b.ts_out != nullptr) reinterpret_cast<uint64_t *>(stage_aux(s) + TILE * 8)[local] = ts;
                    } else if (MODE == MODE_INGEST && a.inplace) {
                        reinterpret_cast<uint32_t *>(stage_aux(s))[ctid] = slot; // the record stays where it is: position = tuple index
                    } else {
                        TileIO<R>::store(buf, local, res);
                        reinterpret_cast<uint32_t *>(stage_aux(s))[local] = slot;
                    }
                } else if (MODE == MODE_INGEST && a.inplace) reinterpret_cast<uint32_t *>(stage_aux(s))[ctid] = INVALID_SLOT;
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&staged[s]);
        }
    } else {
        // ================================= EPILOGUE =================================
        for (uint32_t it = 0;; it++) {
            const uint32_t s = it % STAGES, par = (it / STAGES) & 1u;
            unsigned char *buf = stage_buf(s);
            mbar_wait(&staged[s], par);
            const StageMeta m = meta[s];
            if (m.tile == TILE_SENTINEL) break;
            DevBatch b;
            if (a.batches == nullptr) b = a.one; else b = a.batches[m.batch];
            const uint32_t t = m.tile, tile_count = m.count;
            uint32_t excl = 0;
            if (MODE == MODE_INGEST && a.sparse) {
                excl = t * TILE; // the tile's own region
                if (lane == 0 && t == 0 && a.ff.n_trig != nullptr) { *a.ff.n_trig = 0; *a.ff.n_heavy = 0; } // per-segment lists filled by the update kernels
            } else if constexpr (MODE != MODE_MAP) {
                const uint32_t chain_begin = (MODE == MODE_FILTER) ? b.tile_begin : 0u;
                if (t != chain_begin) {
                    int64_t idx = static_cast<int64_t>(t) - 1;
                    while (true) {
                        const int64_t my = idx - lane;
                        uint64_t w = pack_state(a.epoch, ST_PREFIX, 0); // virtual terminator below the chain start
                        bool valid;
                        do {
                            valid = true;
                            if (my >= static_cast<int64_t>(chain_begin)) {
                                w = ld_relaxed_u64(&a.tile_state[my]);
                                valid = ((w >> 34) == (a.epoch & 0x3fffffffu)) && (((w >> 32) & 3u) != 0);
                            }
                        } while (!__all_sync(FULL, valid));
                        const uint32_t status = static_cast<uint32_t>(w >> 32) & 3u;
                        const uint32_t pmask = __ballot_sync(FULL, status == ST_PREFIX);
                        const uint32_t firstp = pmask ? (__ffs(pmask) - 1) : 32;
                        uint32_t v = (lane <= firstp) ? static_cast<uint32_t>(w) : 0u;
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
                        excl += v;
                        if (pmask) break;
                        idx -= 32;
                    }
                }
                if (lane == 0) {
                    st_relaxed_u64(&a.tile_state[t], pack_state(a.epoch, ST_PREFIX, excl + tile_count));
                    if constexpr (MODE == MODE_FILTER) {
                        const uint32_t last_tile = b.tile_begin + (b.n + TILE - 1) / TILE - 1;
                        if (t == last_tile && b.n_out != nullptr) *b.n_out = excl + tile_count;
                    } else {
                        if (t == 0) { *a.ff.n_trig = 0; *a.ff.n_heavy = 0; } // per-segment lists filled by the update kernels
                        if (t == b.tile_begin) a.batch_off[m.batch] = excl;
                        if (t == a.num_tiles - 1) { a.batch_off[a.nbatches] = excl + tile_count; *a.n_total = excl + tile_count; }
                    }
                }
            }
            // ---- write out ---------------------------------------------------------------------------------------
            unsigned char *dst;
            uint32_t bytes;
            if constexpr (MODE == MODE_MAP) { dst = b.out + static_cast<size_t>(m.first) * TB; bytes = m.cnt * TB; }
            else if constexpr (MODE == MODE_FILTER) { dst = b.out + static_cast<size_t>(excl) * TB; bytes = tile_count * TB; }
            else { dst = a.lifted + static_cast<size_t>(excl) * RB; bytes = a.inplace ? 0u : tile_count * RB; }
            if (MODE == MODE_MAP && (m.flags & TF_SWZ)) {
                if (lane == 0) tma_store_2d(&tmap, 0, static_cast<int32_t>((reinterpret_cast<uint64_t>(dst) - a.tmap_base) >> 6), buf);
            } else if (bulk_ok(dst, bytes)) {
                if (lane == 0 && bytes) {
                    if ((MODE == MODE_INGEST) && a.l2_hints) bulk_s2g_hint(dst, buf, bytes, l2_policy_evict_last());
                    else bulk_s2g(dst, buf, bytes);
                }
            } else {
                uint64_t *d8 = reinterpret_cast<uint64_t *>(dst);
                const uint64_t *s8 = reinterpret_cast<const uint64_t *>(buf);
                for (uint32_t w = lane; w < bytes / 8; w += 32) d8[w] = s8[w];
            }
            if constexpr (MODE == MODE_FILTER) {
                if (b.ts_out != nullptr) {
                    const uint64_t *sts = reinterpret_cast<const uint64_t *>(stage_aux(s) + TILE * 8);
                    for (uint32_t i = lane; i < tile_count; i += 32) b.ts_out[excl + i] = sts[i];
                }
            }
            if constexpr (MODE == MODE_INGEST) {
                const uint32_t *ssl = reinterpret_cast<const uint32_t *>(stage_aux(s));
                if (a.sparse) for (uint32_t i = lane; i < TILE; i += 32) a.slots[excl + i] = i < (a.inplace ? m.cnt : tile_count) ? ssl[i] : INVALID_SLOT;
                else for (uint32_t i = lane; i < tile_count; i += 32) a.slots[excl + i] = ssl[i];
            }
            __syncwarp();
            if (lane == 0) {
                bulk_commit();
                bulk_wait_read<0>();   // the store has finished READING shared memory: the stage can be refilled
                mbar_arrive(&empty[s]);
            }
        }
        if (lane == 0) bulk_wait_all<0>();
    }
}

// ------------------------------------------------------------------------------------------------------
// record helpers (R = result_t): vectorised global load/store and warp shuffles of whole records
// ------------------------------------------------------------------------------------------------------
template <class R>
__device__ __forceinline__ void ld_rec(const unsigned char *p, R &r)
{
    if constexpr (sizeof(R) % 16 == 0) {
        const uint4 *s = reinterpret_cast<const uint4 *>(p);
        uint4 *d = reinterpret_cast<uint4 *>(&r);
#pragma unroll
        for (uint32_t k = 0; k < sizeof(R) / 16; k++) d[k] = s[k];
    } else {
        const uint64_t *s = reinterpret_cast<const uint64_t *>(p);
        uint64_t *d = reinterpret_cast<uint64_t *>(&r);
#pragma unroll
        for (uint32_t k = 0; k < sizeof(R) / 8; k++) d[k] = s[k];
    }
}
template <class R>
__device__ __forceinline__ void st_rec(unsigned char *p, const R &r)
{
    if constexpr (sizeof(R) % 16 == 0) {
        uint4 *d = reinterpret_cast<uint4 *>(p);
        const uint4 *s = reinterpret_cast<const uint4 *>(&r);
#pragma unroll
        for (uint32_t k = 0; k < sizeof(R) / 16; k++) d[k] = s[k];
    } else {
        uint64_t *d = reinterpret_cast<uint64_t *>(p);
        const uint64_t *s = reinterpret_cast<const uint64_t *>(&r);
#pragma unroll
        for (uint32_t k = 0; k < sizeof(R) / 8; k++) d[k] = s[k];
    }
}
template <class R>
__device__ __forceinline__ R shfl_down_rec(const R &r, uint32_t delta)
{
    alignas(16) R o;
    const uint32_t *s = reinterpret_cast<const uint32_t *>(&r);
    uint32_t *d = reinterpret_cast<uint32_t *>(&o);
#pragma unroll
    for (uint32_t k = 0; k < sizeof(R) / 4; k++) d[k] = __shfl_down_sync(FULL, s[k], delta);
    return o;
}
template <class R>
__device__ __forceinline__ R shfl_rec(const R &r, uint32_t src)
{
    alignas(16) R o;
    const uint32_t *s = reinterpret_cast<const uint32_t *>(&r);
    uint32_t *d = reinterpret_cast<uint32_t *>(&o);
#pragma unroll
    for (uint32_t k = 0; k < sizeof(R) / 4; k++) d[k] = __shfl_sync(FULL, s[k], src);
    return o;
}

// exclusive scan of `total` uint32 counters (in may alias out), one CTA of 1024 threads
static __global__ void __launch_bounds__(1024) k_scan_u32(const uint32_t *in, uint32_t *out, uint32_t total, uint32_t *sum_out)
{
    __shared__ uint32_t warp_sums[32];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t per = (total + 1023) / 1024;
    const uint32_t begin = min(tid * per, total), end = min(begin + per, total);
    uint32_t sum = 0;
    for (uint32_t i = begin; i < end; i++) sum += in[i];
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += v; }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = warp_sums[lane], wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(FULL, wi, o); if (lane >= o) wi += v; }
        warp_sums[lane] = wi - w; // exclusive
        if (lane == 31 && sum_out != nullptr) *sum_out = wi;
    }
    __syncthreads();
    uint32_t run = warp_sums[warp] + incl - sum;
    for (uint32_t i = begin; i < end; i++) { const uint32_t v = in[i]; out[i] = run; run += v; }
}


// ------------------------------------------------------------------------------------------------------
// Onesweep-style stable LSD radix pass (8-bit digits): ONE kernel per pass.
//   k_radix_ghist     digit histograms of every pass in one read of the keys   ghist[pass][256]
//   k_onesweep_pass   per tile (dynamic ticket order): stable in-tile ranks (warp match_any, warps in index order),
//                     per-digit decoupled look-back over the earlier tiles (thread d owns digit d, epoch-tagged
//                     64-bit state words), keys/values regrouped by digit in shared memory and written coalesced.
// ctl layout (uint32): [pass][256] histograms, then [pass] ticket counters; the host clears it once per sort.
// ------------------------------------------------------------------------------------------------------
constexpr int OS_THREADS = 256;
constexpr int OS_MAX_PASSES = 8;
// elements per thread (ITEMS): <= 16 for 32-bit keys, <= 8 for 64-bit keys (static smem <= 48 KB)
template <class K> struct OsCfg { static constexpr int MAX_ITEMS = sizeof(K) == 4 ? 16 : 8; };

template <class K>
__global__ void __launch_bounds__(256) k_radix_ghist(const K *__restrict__ keys, const uint32_t *__restrict__ n_ptr, uint32_t n_host,
                                                     uint32_t passes, uint32_t *__restrict__ ctl, uint32_t base_shift)
{
    __shared__ uint32_t h[OS_MAX_PASSES * 256];
    const uint32_t n = n_ptr ? *n_ptr : n_host;
    for (uint32_t i = threadIdx.x; i < passes * 256; i += blockDim.x) h[i] = 0;
    __syncthreads();
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const K k = keys[i];
        for (uint32_t p = 0; p < passes; p++) atomicAdd(&h[p * 256 + (static_cast<uint32_t>(k >> (base_shift + 8 * p)) & 255u)], 1u);
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < passes * 256; i += blockDim.x) if (h[i]) atomicAdd(&ctl[i], h[i]);
}

template <class K, int OS_ITEMS>
__global__ void __launch_bounds__(OS_THREADS) k_onesweep_pass(const K *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                                                              K *__restrict__ keys_out, uint32_t *__restrict__ vals_out,
                                                              const uint32_t *__restrict__ n_ptr, uint32_t n_host, uint32_t pass,
                                                              uint32_t passes, uint32_t *__restrict__ ctl,
                                                              uint64_t *__restrict__ tile_state, uint32_t epoch,
                                                              const unsigned char *__restrict__ payload_in,
                                                              unsigned char *__restrict__ payload_out, uint32_t payload_bytes,
                                                              uint32_t *__restrict__ seg_first, uint32_t seg_first_n,
                                                              uint32_t base_shift)
{
    __shared__ uint32_t cntw[OS_THREADS / 32][256]; // per-warp digit counts -> exclusive offsets over the warps
    __shared__ uint32_t dig_off[256];               // exclusive offset of each digit inside the tile
    __shared__ uint32_t bin_base[256];              // global position of the tile's first element of each digit
    __shared__ uint32_t wsum[OS_THREADS / 32];
    __shared__ uint32_t s_tile;
    constexpr int OS_TILE = OS_THREADS * OS_ITEMS;
    __shared__ K skeys[OS_TILE];
    __shared__ uint32_t svals[OS_TILE];

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t n = n_ptr ? *n_ptr : n_host;
    const uint32_t num_tiles = (n + OS_TILE - 1) / OS_TILE;
    const uint32_t shift = base_shift + 8 * pass;
    if (tid == 0) s_tile = atomicAdd(&ctl[passes * 256 + pass], 1u);
#pragma unroll
    for (int w = 0; w < OS_THREADS / 32; w++) cntw[w][tid] = 0;
    __syncthreads();
    const uint32_t tile = s_tile;
    if (tile >= num_tiles) return;
    const uint32_t start = tile * OS_TILE;

    // ---- stable in-tile ranks: warp w owns [start + w*512, +512), 32 consecutive elements per round ------------
    K k[OS_ITEMS];
    uint32_t rk[OS_ITEMS];
#pragma unroll
    for (int r = 0; r < OS_ITEMS; r++) {
        const uint32_t idx = start + warp * (32 * OS_ITEMS) + r * 32 + lane;
        const bool valid = idx < n;
        k[r] = valid ? keys_in[idx] : K(0);
        const uint32_t d = valid ? (static_cast<uint32_t>(k[r] >> shift) & 255u) : 256u;
        const uint32_t mask = __match_any_sync(FULL, d);
        rk[r] = valid ? (cntw[warp][d] + __popc(mask & lanemask_lt())) : 0u;
        __syncwarp();
        if (valid && lane == static_cast<uint32_t>(__ffs(mask) - 1)) cntw[warp][d] += __popc(mask);
        __syncwarp();
    }
    __syncthreads();

    // ---- digit `tid`: tile total, exclusive offsets over warps, publish, look back -----------------------------------
    uint32_t total = 0;
#pragma unroll
    for (int w = 0; w < OS_THREADS / 32; w++) { const uint32_t c = cntw[w][tid]; cntw[w][tid] = total; total += c; }
    uint64_t *my_state = tile_state + static_cast<size_t>(tile) * 256 + tid;
    st_relaxed_u64(my_state, pack_state(epoch, tile == 0 ? ST_PREFIX : ST_AGG, total));
    // global base of digit tid = exclusive scan of the pass histogram over the digits
    const uint32_t gcount = ctl[pass * 256 + tid];
    uint32_t incl = gcount;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += v; }
    if (lane == 31) wsum[warp] = incl;
    // exclusive offset of the digit inside the tile (same scan over `total`)
    uint32_t tincl = total;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(FULL, tincl, o); if (lane >= o) tincl += v; }
    __shared__ uint32_t twsum[OS_THREADS / 32];
    if (lane == 31) twsum[warp] = tincl;
    __syncthreads();
    uint32_t gbase = incl - gcount, tbase = tincl - total;
#pragma unroll
    for (uint32_t w = 0; w < OS_THREADS / 32; w++) if (w < warp) { gbase += wsum[w]; tbase += twsum[w]; }
    uint32_t excl = 0;
    if (tile > 0) { // thread d walks digit d's chain back, 8 independent loads per step
        int64_t t2 = static_cast<int64_t>(tile) - 1;
        bool found = false;
        while (!found) {
            uint64_t w[8];
#pragma unroll
            for (int q = 0; q < 8; q++) w[q] = (t2 - q >= 0) ? ld_relaxed_u64(tile_state + static_cast<size_t>(t2 - q) * 256 + tid) : 0ull;
#pragma unroll
            for (int q = 0; q < 8; q++) {
                if (found || t2 - q < 0) continue;
                if ((w[q] >> 34) != (epoch & 0x3fffffffu) || ((w[q] >> 32) & 3u) == 0) { t2 -= q; goto next_round; } // not published yet: retry from here
                excl += static_cast<uint32_t>(w[q]);
                if (((w[q] >> 32) & 3u) == ST_PREFIX) found = true;
            }
            t2 -= 8;
        next_round:;
        }
        st_relaxed_u64(my_state, pack_state(epoch, ST_PREFIX, excl + total));
    }
    dig_off[tid] = tbase;
    bin_base[tid] = gbase + excl;
    __syncthreads();

    // ---- regroup by digit in shared memory, then coalesced writes ---------------------------------------------------
#pragma unroll
    for (int r = 0; r < OS_ITEMS; r++) {
        const uint32_t idx = start + warp * (32 * OS_ITEMS) + r * 32 + lane;
        if (idx < n) {
            const uint32_t d = static_cast<uint32_t>(k[r] >> shift) & 255u;
            const uint32_t lp = dig_off[d] + cntw[warp][d] + rk[r];
            skeys[lp] = k[r];
            svals[lp] = vals_in ? vals_in[idx] : idx;
        }
    }
    __syncthreads();
    const uint32_t cnt = min(static_cast<uint32_t>(OS_TILE), n - start);
    for (uint32_t i = tid; i < cnt; i += OS_THREADS) {
        const K kk = skeys[i];
        const uint32_t d = static_cast<uint32_t>(kk >> shift) & 255u;
        const uint32_t dst = bin_base[d] + (i - dig_off[d]);
        keys_out[dst] = kk;
        const uint32_t v = svals[i];
        vals_out[dst] = v;
        // last pass of the window operator's sort: first sorted position of every key (entries start at 0xffffffff)
        if (seg_first != nullptr && static_cast<uint64_t>(kk) < seg_first_n) atomicMin(&seg_first[static_cast<uint32_t>(kk)], dst);
        if (payload_out != nullptr) { // the last pass also moves the records: out[dst] = in[value]
            if ((payload_bytes & 15u) == 0) {
                const uint4 *src = reinterpret_cast<const uint4 *>(payload_in + static_cast<size_t>(v) * payload_bytes);
                uint4 *dstp = reinterpret_cast<uint4 *>(payload_out + static_cast<size_t>(dst) * payload_bytes);
                for (uint32_t q = 0; q < payload_bytes / 16; q++) dstp[q] = src[q];
            } else {
                const uint64_t *src = reinterpret_cast<const uint64_t *>(payload_in + static_cast<size_t>(v) * payload_bytes);
                uint64_t *dstp = reinterpret_cast<uint64_t *>(payload_out + static_cast<size_t>(dst) * payload_bytes);
                for (uint32_t q = 0; q < payload_bytes / 8; q++) dstp[q] = src[q];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// k_slots_inplace: the streaming pass of a pass-through program whose records already sit at their tile positions
// (TileArgs::inplace): nothing is staged or copied, so a plain grid-stride kernel replaces k_tile_pass -- key -> slot (or the
// caller's slot), INVALID_SLOT padding, the digit counts of the wide partition (shared-memory counts, one flush per CTA).
// 32-byte records make 8-KB tiles, too small to amortise the tile pass's per-tile machinery (measured 1.2 TB/s there).
// ------------------------------------------------------------------------------------------------------
template <class P>
__global__ void __launch_bounds__(256) k_slots_inplace(const TileArgs a, const typename P::params_t prm)
{
    using T = typename P::tuple_t;
    __shared__ uint32_t s_h[1024];
    const uint32_t tid = threadIdx.x;
    const uint32_t npos = a.num_tiles * TILE;
    if (blockIdx.x == 0 && tid == 0 && a.ff.n_trig != nullptr) { *a.ff.n_trig = 0; *a.ff.n_heavy = 0; } // per-segment lists of the update kernels
    auto slot_at = [&](uint32_t p) -> uint32_t {
        const uint32_t t = p / TILE;
        uint32_t lo = 0, hi = a.nbatches - 1; // last batch with tile_begin <= t
        while (lo < hi) { const uint32_t mid = (lo + hi + 1) >> 1; if (a.batches[mid].tile_begin <= t) lo = mid; else hi = mid - 1; }
        const DevBatch &b = a.batches[lo];
        const uint32_t local = p - b.tile_begin * TILE;
        uint32_t slot = INVALID_SLOT;
        if (local < b.n) {
            if (a.ext_slots != nullptr) { slot = a.ext_slots[p]; if (slot >= a.ff.max_keys) slot = INVALID_SLOT; }
            else {
                const T *rec = reinterpret_cast<const T *>(b.tuples + static_cast<size_t>(local) * sizeof(T));
                slot = slot_of_key(a.ff, P::key(*rec, prm));
            }
            if (slot != INVALID_SLOT && a.count_keys) atomicAdd(&a.ff.seg_cnt[slot], 1u);
        }
        return slot;
    };
    if (a.wide_h16 != nullptr) {
        // a CTA owns whole wide tiles (4096 positions): digit counts in shared memory (two 16-bit counters per word), one row per wide tile,
        // and -- TileArgs::pack_rank -- every slot leaves with the count its digit had when it was counted (k_wide_scatter_ranked)
        const uint32_t nwide = (npos + OSW_TILE_POS - 1) / OSW_TILE_POS;
        for (uint32_t wt = blockIdx.x; wt < nwide; wt += gridDim.x) {
            for (uint32_t i = tid; i < 512; i += blockDim.x) s_h[i] = 0;
            __syncthreads();
            for (uint32_t p = wt * OSW_TILE_POS + tid; p < min(npos, (wt + 1) * OSW_TILE_POS); p += blockDim.x) {
                uint32_t slot = slot_at(p);
                if (slot != INVALID_SLOT) {
                    const uint32_t d = (slot >> a.sort_shift) & 1023u;
                    const uint32_t before = atomicAdd(&s_h[d >> 1], 1u << ((d & 1u) * 16u));
                    if (a.pack_rank) slot |= ((before >> ((d & 1u) * 16u)) & 0xffffu) << 16;
                }
                a.slots[p] = slot;
            }
            __syncthreads();
            for (uint32_t i = tid; i < 512; i += blockDim.x) reinterpret_cast<uint32_t *>(a.wide_h16 + static_cast<size_t>(wt) * 1024u)[i] = s_h[i];
            __syncthreads();
        }
        return;
    }
    const bool hist = a.sort_ctl != nullptr;
    if (hist) for (uint32_t i = tid; i < 1024; i += blockDim.x) s_h[i] = 0;
    __syncthreads();
    const uint32_t dmask = (1u << a.sort_dbits) - 1u;
    for (uint32_t p = blockIdx.x * blockDim.x + tid; p < npos; p += gridDim.x * blockDim.x) {
        const uint32_t slot = slot_at(p);
        if (slot != INVALID_SLOT) {
            if (hist) atomicAdd(&s_h[(slot >> a.sort_shift) & dmask], 1u);
            if (a.wide_h32 != nullptr) atomicAdd(&a.wide_h32[static_cast<size_t>(p / OSW_TILE_POS) * 1024u + ((slot >> a.sort_shift) & 1023u)], 1u);
        }
        a.slots[p] = slot;
    }
    if (hist) {
        __syncthreads();
        for (uint32_t i = tid; i < (1u << a.sort_dbits) && i < 1024; i += blockDim.x) { const uint32_t c = s_h[i]; if (c) atomicAdd(&a.sort_ctl[i], c); }
    }
}

// ------------------------------------------------------------------------------------------------------
// Wide partition: ONE stable pass on a 10-bit digit (1024 bins) without a chained scan -- with 1024 bins a tile holds
// only a few elements per bin, so the look-back chains of k_onesweep_pass are long and cheap to avoid:
//   k_wide_tile_hist  per-tile digit counts H[tile][1024] (16-bit) + their sums over chunks of 2^chunk_shift tiles
//                     C[chunk][1024] (+ the global counts ctl[1024] unless the producer of the keys already made them)
//   k_wide_scatter    rank inside the tile (warp match_any, warps in index order), base of (tile, digit) =
//                     exclusive scan of ctl over the digits + C rows of the earlier chunks + H rows of the earlier
//                     tiles of the own chunk; elements go straight to their final position
// The window operator uses it to split a segment's (slot, arrival position) pairs into 1024 buckets of consecutive
// slots (k_ffat_update_buckets finishes the grouping inside each bucket).
// ------------------------------------------------------------------------------------------------------
#ifndef WFB_OSW_MINBLOCKS
#define WFB_OSW_MINBLOCKS 5
#endif
constexpr uint32_t OSW_BITS = 10, OSW_DIGITS = 1u << OSW_BITS;
constexpr uint32_t OSW_THREADS = 256, OSW_ITEMS = 16, OSW_TILE = OSW_THREADS * OSW_ITEMS; // 4096 elements per tile
static_assert(OSW_TILE == OSW_TILE_POS, "the tile pass files its digit counts per wide tile");

template <class K>
__global__ void __launch_bounds__(OSW_THREADS) k_wide_tile_hist(const K *__restrict__ keys, const uint32_t *__restrict__ n_ptr, uint32_t n_host,
                                                                uint32_t shift, uint32_t chunk_shift, uint16_t *__restrict__ H,
                                                                uint32_t *__restrict__ C, uint32_t *__restrict__ ctl_counts, uint32_t skip_invalid)
{
    static_assert(OSW_THREADS * 4 == OSW_DIGITS, "four digits per thread");
    __shared__ __align__(16) uint32_t h[OSW_DIGITS];
    const uint32_t tid = threadIdx.x, tile = blockIdx.x;
    const uint32_t n = n_ptr ? *n_ptr : n_host;
    const uint32_t start = tile * OSW_TILE;
    if (start >= n) return;
    reinterpret_cast<uint4 *>(h)[tid] = make_uint4(0, 0, 0, 0);
    __syncthreads();
#pragma unroll
    for (uint32_t r = 0; r < OSW_ITEMS; r++) {
        const uint32_t idx = start + r * OSW_THREADS + tid;
        if (idx < n) { const K kk = keys[idx]; if (!(skip_invalid && kk == static_cast<K>(~K(0)))) atomicAdd(&h[static_cast<uint32_t>(kk >> shift) & (OSW_DIGITS - 1u)], 1u); }
    }
    __syncthreads();
    const uint4 c = reinterpret_cast<const uint4 *>(h)[tid];
    reinterpret_cast<ushort4 *>(H + static_cast<size_t>(tile) * OSW_DIGITS)[tid] =
        make_ushort4(static_cast<uint16_t>(c.x), static_cast<uint16_t>(c.y), static_cast<uint16_t>(c.z), static_cast<uint16_t>(c.w));
    uint32_t *crow = C + static_cast<size_t>(tile >> chunk_shift) * OSW_DIGITS + tid * 4;
    if (c.x) atomicAdd(crow + 0, c.x);
    if (c.y) atomicAdd(crow + 1, c.y);
    if (c.z) atomicAdd(crow + 2, c.z);
    if (c.w) atomicAdd(crow + 3, c.w);
    if (ctl_counts != nullptr) {
        if (c.x) atomicAdd(ctl_counts + tid * 4 + 0, c.x);
        if (c.y) atomicAdd(ctl_counts + tid * 4 + 1, c.y);
        if (c.z) atomicAdd(ctl_counts + tid * 4 + 2, c.z);
        if (c.w) atomicAdd(ctl_counts + tid * 4 + 3, c.w);
    }
}

// C[chunk][digit] = sum of the 32-bit per-tile rows of the chunk (when the tile pass filled them: no counting pass)
static __global__ void __launch_bounds__(OSW_THREADS) k_wide_chunk_sums(const uint32_t *__restrict__ H32, uint32_t tiles, uint32_t chunk_shift, uint32_t *__restrict__ C)
{
    const uint32_t chunk = blockIdx.x, tid = threadIdx.x;
    const uint32_t t0 = chunk << chunk_shift, t1 = min(tiles, t0 + (1u << chunk_shift));
    uint4 acc = make_uint4(0, 0, 0, 0);
    const uint4 *row = reinterpret_cast<const uint4 *>(H32) + tid;
#pragma unroll 8
    for (uint32_t t = t0; t < t1; t++) { const uint4 v = row[static_cast<size_t>(t) * (OSW_DIGITS / 4)]; acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
    reinterpret_cast<uint4 *>(C + static_cast<size_t>(chunk) * OSW_DIGITS)[tid] = acc;
}

// the same for the 16-bit rows the tile pass files (TileArgs::wide_h16)
static __global__ void __launch_bounds__(OSW_THREADS) k_wide_chunk_sums16(const uint16_t *__restrict__ H, uint32_t tiles, uint32_t chunk_shift, uint32_t *__restrict__ C,
                                                                          uint32_t *__restrict__ ctl_counts = nullptr)
{
    // ctl_counts (optional): the global digit counts are accumulated there as well (callers that do not run k_wide_chunk_scan)
    const uint32_t chunk = blockIdx.x, tid = threadIdx.x;
    const uint32_t t0 = chunk << chunk_shift, t1 = min(tiles, t0 + (1u << chunk_shift));
    uint4 acc = make_uint4(0, 0, 0, 0);
    const ushort4 *row = reinterpret_cast<const ushort4 *>(H) + tid;
#pragma unroll 16
    for (uint32_t t = t0; t < t1; t++) { const ushort4 v = row[static_cast<size_t>(t) * (OSW_DIGITS / 4)]; acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
    reinterpret_cast<uint4 *>(C + static_cast<size_t>(chunk) * OSW_DIGITS)[tid] = acc;
    if (ctl_counts != nullptr) {
        if (acc.x) atomicAdd(ctl_counts + tid * 4 + 0, acc.x);
        if (acc.y) atomicAdd(ctl_counts + tid * 4 + 1, acc.y);
        if (acc.z) atomicAdd(ctl_counts + tid * 4 + 2, acc.z);
        if (acc.w) atomicAdd(ctl_counts + tid * 4 + 3, acc.w);
    }
}

// C[chunk][digit] (chunk sums) -> first output position of (chunk, digit): exclusive scan over the digits of the totals + exclusive
// scan over the chunks, in place; ctl_counts[digit] = total of the digit. One CTA, one thread per digit, every load independent.
static __global__ void __launch_bounds__(OSW_DIGITS) k_wide_chunk_scan(uint32_t *__restrict__ C, uint32_t chunks, uint32_t *__restrict__ ctl_counts)
{
    __shared__ uint32_t wsum[32];
    const uint32_t d = threadIdx.x, lane = d & 31, warp = d >> 5;
    uint32_t total = 0;
#pragma unroll 16
    for (uint32_t c = 0; c < chunks; c++) total += C[static_cast<size_t>(c) * OSW_DIGITS + d];
    ctl_counts[d] = total;
    uint32_t incl = total;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(FULL, incl, o); if (lane >= static_cast<uint32_t>(o)) incl += v; }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const uint32_t w = wsum[lane];
        uint32_t wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(FULL, wi, o); if (lane >= static_cast<uint32_t>(o)) wi += v; }
        wsum[lane] = wi - w;
    }
    __syncthreads();
    uint32_t run = wsum[warp] + incl - total;
#pragma unroll 16
    for (uint32_t c = 0; c < chunks; c++) { uint32_t *p = C + static_cast<size_t>(c) * OSW_DIGITS + d; const uint32_t v = *p; *p = run; run += v; }
}

// The scatter of the wide partition when the tile pass packed a rank with every slot (TileArgs::pack_rank): word = slot | rank << 16,
// rank = any numbering 0 .. count-1 of the survivors of one (wide tile, digit) cell. One CTA per wide tile, everything on chip:
//   1. first output position of every digit for this tile (chunk row of k_wide_chunk_scan + the rows of the earlier tiles of the chunk)
//      and the tile's own cells laid out back to back in shared memory (exclusive scan of the tile's digit counts),
//   2. every item files its position within the tile at cell start + rank -- no ranking rounds, no per-warp counters,
//   3. ARRIVAL order inside a cell (the count windows need every key's items in stream order, and two items of one key may share a
//      cell): an item's place = number of the cell's entries with a smaller position; a cell holds a few entries, read from shared memory,
//   4. one write of (slot, position) per item to its final place.
// Output: keys_out[i] = slot, vals_out[i] = arrival position, stable by (digit, position) -- what k_wide_scatter produces.
// RBYTES != 0: the records travel (payload_in at the arrival positions -> payload_out at the final places; keys_out gets the slots,
// vals_out is not written): the source side of the bucketed multi-GPU exchange.
#ifndef WFB_OSR_MINBLOCKS
#define WFB_OSR_MINBLOCKS 1   // (resident CTAs per SM the pair version is compiled for: 2048 tiles of the bench step are 2 full waves at 7)
#endif
template <int RBYTES>
static __global__ void __launch_bounds__(OSW_THREADS, RBYTES == 0 ? WFB_OSR_MINBLOCKS : 1) k_wide_scatter_ranked(const uint32_t *__restrict__ packed, uint32_t *__restrict__ keys_out,
                                                                            uint32_t *__restrict__ vals_out, uint32_t n, uint32_t shift, uint32_t chunk_shift,
                                                                            const uint16_t *__restrict__ H, const uint32_t *__restrict__ Cx,
                                                                            const unsigned char *__restrict__ payload_in, unsigned char *__restrict__ payload_out)
{
    constexpr uint32_t NW = OSW_THREADS / 32;
    __shared__ __align__(16) uint32_t bin_base[OSW_DIGITS];
    __shared__ __align__(8) uint16_t cnt_row[OSW_DIGITS], cell_start[OSW_DIGITS];
    constexpr uint32_t LPOS = OSW_TILE + 3u * OSW_DIGITS;  // every cell padded to a multiple of four entries (8-byte loads in the repair loop)
    constexpr uint32_t LPAD = 0x0fffu;                     // padding entry: never below a position of the tile (positions are < 4096)
    static_assert(OSW_TILE <= 4096 && LPOS % (2 * OSW_THREADS) == 0, "16-bit packed position compare");
    __shared__ __align__(8) uint16_t lpos[LPOS];
    __shared__ uint32_t wsum[NW];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, tile = blockIdx.x;
    const uint32_t start = tile * OSW_TILE;
    if (start >= n) return;
    uint32_t w[OSW_ITEMS];
#pragma unroll
    for (uint32_t r = 0; r < OSW_ITEMS; r++) { const uint32_t idx = start + r * OSW_THREADS + tid; w[r] = idx < n ? packed[idx] : INVALID_SLOT; }
    if constexpr (RBYTES != 0) { // the records this CTA will move: on their way to L2 while the offsets are worked out below
#pragma unroll
        for (uint32_t r = 0; r < OSW_ITEMS; r++)
            if (w[r] != INVALID_SLOT) asm volatile("prefetch.global.L2 [%0];" ::"l"(payload_in + static_cast<size_t>(start + r * OSW_THREADS + tid) * RBYTES));
    }
    {
        const uint32_t chunk = tile >> chunk_shift;
        uint4 acc = reinterpret_cast<const uint4 *>(Cx)[static_cast<size_t>(chunk) * (OSW_DIGITS / 4) + tid];
        const ushort4 *hrow = reinterpret_cast<const ushort4 *>(H) + tid;
#pragma unroll 16
        for (uint32_t t = chunk << chunk_shift; t < tile; t++) { const ushort4 v = hrow[static_cast<size_t>(t) * (OSW_DIGITS / 4)]; acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
        reinterpret_cast<uint4 *>(bin_base)[tid] = acc;
        const ushort4 c = hrow[static_cast<size_t>(tile) * (OSW_DIGITS / 4)];
        reinterpret_cast<ushort4 *>(cnt_row)[tid] = c;
        // the tile's cells back to back: exclusive scan of its 1024 digit counts (thread tid owns digits 4 tid .. 4 tid + 3)
        const uint32_t p0 = (c.x + 3u) & ~3u, p1 = (c.y + 3u) & ~3u, p2 = (c.z + 3u) & ~3u, p3 = (c.w + 3u) & ~3u; // padded cell sizes
        const uint32_t sum = p0 + p1 + p2 + p3;
        uint32_t incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(FULL, incl, o); if (lane >= static_cast<uint32_t>(o)) incl += v; }
        if (lane == 31) wsum[warp] = incl;
#pragma unroll
        for (uint32_t i = 0; i < LPOS / (2 * OSW_THREADS); i++) reinterpret_cast<uint32_t *>(lpos)[i * OSW_THREADS + tid] = LPAD | (LPAD << 16);
        __syncthreads();
        uint32_t base = incl - sum;
#pragma unroll
        for (uint32_t q = 0; q < NW; q++) if (q < warp) base += wsum[q];
        reinterpret_cast<ushort4 *>(cell_start)[tid] = make_ushort4(static_cast<uint16_t>(base), static_cast<uint16_t>(base + p0), static_cast<uint16_t>(base + p0 + p1),
                                                                     static_cast<uint16_t>(base + p0 + p1 + p2));
    }
    __syncthreads();
#pragma unroll
    for (uint32_t r = 0; r < OSW_ITEMS; r++) {
        if (w[r] != INVALID_SLOT) {
            const uint32_t d = ((w[r] & 0xffffu) >> shift) & (OSW_DIGITS - 1u);
            lpos[cell_start[d] + (w[r] >> 16)] = static_cast<uint16_t>(r * OSW_THREADS + tid);
        }
    }
    __syncthreads();
#pragma unroll 4
    for (uint32_t r = 0; r < OSW_ITEMS; r++) {
        if (w[r] != INVALID_SLOT) {
            const uint32_t slot = w[r] & 0xffffu, d = (slot >> shift) & (OSW_DIGITS - 1u);
            const uint32_t cnt = cnt_row[d], cs = cell_start[d], mine = r * OSW_THREADS + tid;
            // entries of the cell below `mine`, four per load: per 16-bit half, bit 15 of (0x8000 + mine - 1 - entry) says entry < mine
            // (all values are < 4096, so the halves never borrow from each other; padding entries are never below)
            const uint32_t mm = (mine | (mine << 16)) + 0x7fff7fffu;
            const uint2 *cell = reinterpret_cast<const uint2 *>(lpos + cs);
            uint32_t less = 0;
            for (uint32_t e = 0; e < cnt; e += 4) { const uint2 v = cell[e >> 2]; less += __popc((mm - v.x) & 0x80008000u) + __popc((mm - v.y) & 0x80008000u); }
            const uint32_t dst = bin_base[d] + less;
            keys_out[dst] = slot;
            if constexpr (RBYTES == 0) vals_out[dst] = start + mine;
            else {
                static_assert(RBYTES % 8 == 0, "record size");
                using W = typename std::conditional<RBYTES % 16 == 0, uint4, uint2>::type;
                const W *src = reinterpret_cast<const W *>(payload_in + static_cast<size_t>(start + mine) * RBYTES);
                W *dstp = reinterpret_cast<W *>(payload_out + static_cast<size_t>(dst) * RBYTES);
                W v[RBYTES / sizeof(W)];
#pragma unroll
                for (uint32_t q = 0; q < RBYTES / sizeof(W); q++) v[q] = src[q];
#pragma unroll
                for (uint32_t q = 0; q < RBYTES / sizeof(W); q++) dstp[q] = v[q];
            }
        }
    }
}

// bucketed exchange, source side: records per destination = sums of the bins [d * bps, (d + 1) * bps) of the partition
// (send_meta != nullptr: also the (count, watermark) pair every destination is sent ahead of the records)
static __global__ void k_shard_bin_counts(const uint32_t *__restrict__ bin_counts, uint32_t nshards, uint32_t bps, uint32_t *__restrict__ counts_out,
                                          uint64_t *__restrict__ send_meta, uint64_t watermark)
{
    const uint32_t d = threadIdx.x >> 5, lane = threadIdx.x & 31; // one warp per destination
    uint32_t c = 0;
    if (d < nshards) for (uint32_t b = lane; b < bps; b += 32) c += bin_counts[d * bps + b];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(FULL, c, o);
    if (lane == 0 && d < MAX_SHARDS) counts_out[d] = d < nshards ? c : 0u; // ([MAX_SHARDS]: error flags, set by the tile pass)
    if (lane == 0 && d < nshards && send_meta != nullptr) { send_meta[2 * d] = c; send_meta[2 * d + 1] = watermark; }
}

// bucketed exchange, destination side: source s delivered, for every COARSE bucket b of this GPU's slot space (bps of them: the
// source's 1024 bins are shared by all destinations), a run of cnt[s][b] records in arrival order -- the runs of one source back to
// back from recv position off[s]. The update kernel wants all 1024 CTAs busy, so every coarse bucket is split into nsub = 1024 / bps
// sub-buckets by slot: the items of sub-bucket (b, j) are, source after source (= global stream order), the items of run (s, b) whose
// slot falls into j, in arrival order. Three small kernels write them as the (slot, position) lists + sizes k_ffat_update_buckets
// consumes: count per (b, j, s) | exclusive scan in that order | stable split of every run.
struct MgRuns { uint32_t off[MAX_SHARDS + 1]; };
constexpr uint32_t MG_THREADS = 256;
__device__ __forceinline__ uint32_t mg_run_start(const uint32_t *__restrict__ cnt, uint32_t s, uint32_t bps, uint32_t b, uint32_t *sh)
{   // records of source s in the coarse buckets before b (block-wide sum; sh: 8 words of shared memory)
    uint32_t c = 0;
    for (uint32_t q = threadIdx.x; q < b; q += MG_THREADS) c += cnt[s * bps + q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(FULL, c, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = c;
    __syncthreads();
    uint32_t t = 0;
#pragma unroll
    for (uint32_t w = 0; w < MG_THREADS / 32; w++) t += sh[w];
    __syncthreads();
    return t;
}
static __global__ void __launch_bounds__(MG_THREADS) k_mg_count(const uint32_t *__restrict__ cnt, uint32_t nsrc, uint32_t bps, const MgRuns runs,
                                                                const uint32_t *__restrict__ recv_slots, uint32_t slot_mask, uint32_t shift2, uint32_t nsub,
                                                                uint32_t *__restrict__ cnt3, uint32_t *__restrict__ run_starts,
                                                                uint32_t *__restrict__ n_trig, uint32_t *__restrict__ n_heavy)
{
    __shared__ uint32_t sh[8], c[MAX_SHARDS];
    const uint32_t b = blockIdx.x, s = blockIdx.y, tid = threadIdx.x, lane = tid & 31;
    if (b == 0 && s == 0 && tid == 0) { *n_trig = 0; *n_heavy = 0; } // per-segment lists filled by the update kernel
    if (tid < MAX_SHARDS) c[tid] = 0;
    const uint32_t rs = mg_run_start(cnt, s, bps, b, sh), m = cnt[s * bps + b];
    if (tid == 0) run_starts[s * bps + b] = rs;
    const uint32_t *sl = recv_slots + runs.off[s] + rs;
    for (uint32_t i0 = 0; i0 < m; i0 += MG_THREADS) {
        const uint32_t i = i0 + tid;
        const uint32_t j = i < m ? ((sl[i] & slot_mask) >> shift2) & (nsub - 1u) : nsub;
        for (uint32_t jj = 0; jj < nsub; jj++) { const uint32_t bal = __ballot_sync(FULL, j == jj); if (lane == 0 && bal) atomicAdd(&c[jj], __popc(bal)); }
    }
    __syncthreads();
    if (tid < nsub) cnt3[(b * nsub + tid) * nsrc + s] = c[tid];
}
// exclusive scan of cnt3 in (bucket, sub-bucket, source) order + the sub-bucket sizes (one CTA of 1024 threads; n <= 1024 * MAX_SHARDS)
static __global__ void __launch_bounds__(1024) k_mg_scan(const uint32_t *__restrict__ cnt3, uint32_t n, uint32_t nsrc, uint32_t *__restrict__ off3,
                                                         uint32_t *__restrict__ digit_counts)
{
    __shared__ uint32_t wsum[32];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t per = (n + 1023u) / 1024u, lo = tid * per;
    uint32_t v[MAX_SHARDS], sum = 0;
#pragma unroll
    for (uint32_t q = 0; q < MAX_SHARDS; q++) { v[q] = (q < per && lo + q < n) ? cnt3[lo + q] : 0u; sum += v[q]; }
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t x = __shfl_up_sync(FULL, incl, o); if (lane >= static_cast<uint32_t>(o)) incl += x; }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = wsum[lane], wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t x = __shfl_up_sync(FULL, wi, o); if (lane >= static_cast<uint32_t>(o)) wi += x; }
        wsum[lane] = wi - w;
    }
    __syncthreads();
    uint32_t run = wsum[warp] + incl - sum;
#pragma unroll
    for (uint32_t q = 0; q < MAX_SHARDS; q++) if (q < per && lo + q < n) { off3[lo + q] = run; run += v[q]; }
    for (uint32_t d = tid; d * nsrc < n; d += 1024) { uint32_t t = 0; for (uint32_t q = 0; q < nsrc; q++) t += cnt3[d * nsrc + q]; digit_counts[d] = t; }
}
static __global__ void __launch_bounds__(MG_THREADS) k_mg_split(const uint32_t *__restrict__ cnt, uint32_t nsrc, uint32_t bps, const MgRuns runs,
                                                                const uint32_t *__restrict__ recv_slots, uint32_t slot_mask, uint32_t shift2, uint32_t nsub,
                                                                const uint32_t *__restrict__ off3, const uint32_t *__restrict__ run_starts,
                                                                uint32_t *__restrict__ out_slots, uint32_t *__restrict__ out_pos)
{
    __shared__ uint32_t wc[MG_THREADS / 32][MAX_SHARDS], fill[MAX_SHARDS];
    const uint32_t b = blockIdx.x, s = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t m = cnt[s * bps + b], src0 = runs.off[s] + run_starts[s * bps + b];
    if (tid < nsub) fill[tid] = off3[(b * nsub + tid) * nsrc + s];
    __syncthreads();
    for (uint32_t i0 = 0; i0 < m; i0 += MG_THREADS) {
        const uint32_t i = i0 + tid;
        uint32_t slot = 0, j = nsub, before = 0;
        if (i < m) { slot = recv_slots[src0 + i] & slot_mask; j = (slot >> shift2) & (nsub - 1u); }
        for (uint32_t jj = 0; jj < nsub; jj++) {
            const uint32_t bal = __ballot_sync(FULL, j == jj);
            if (lane == 0) wc[warp][jj] = __popc(bal);
            if (j == jj) before = __popc(bal & lanemask_lt());
        }
        __syncthreads();
        if (i < m) {
            uint32_t dst = fill[j] + before;
            for (uint32_t w = 0; w < warp; w++) dst += wc[w][j];
            out_slots[dst] = slot; out_pos[dst] = src0 + i;
        }
        __syncthreads();
        if (tid < nsub) { uint32_t t = 0; for (uint32_t w = 0; w < MG_THREADS / 32; w++) t += wc[w][tid]; fill[tid] += t; }
        __syncthreads();
    }
}

// RBYTES: bytes of the payload record that travels with each element (payload_out[dst] = payload_in[index]); 0 = none,
// -1 = run-time size `payload_bytes` (multiple of 8)
template <class K, int RBYTES>
__global__ void __launch_bounds__(OSW_THREADS, WFB_OSW_MINBLOCKS) k_wide_scatter(const K *__restrict__ keys_in, K *__restrict__ keys_out, uint32_t *__restrict__ vals_out,
                                                              const uint32_t *__restrict__ n_ptr, uint32_t n_host, uint32_t shift,
                                                              uint32_t chunk_shift, const uint16_t *__restrict__ H, const uint32_t *__restrict__ C,
                                                              const uint32_t *__restrict__ ctl_counts,
                                                              const unsigned char *__restrict__ payload_in, unsigned char *__restrict__ payload_out,
                                                              uint32_t payload_bytes, uint32_t skip_invalid, uint32_t region_stride,
                                                              const uint32_t *__restrict__ H32, uint32_t cx)
{
    // cx != 0: C holds the first output position of every (chunk, digit) already (k_wide_chunk_scan); ctl_counts is not read
    // H32 != nullptr: the per-tile counts are 32-bit rows filled by the producer of the keys (the tile pass) instead of H
    // region_stride != 0: bin d starts at d * region_stride (fixed-capacity regions; elements beyond the capacity are dropped)
    constexpr uint32_t NW = OSW_THREADS / 32;
    __shared__ __align__(16) uint16_t cntw[NW][OSW_DIGITS]; // per-warp digit counts -> exclusive offsets over the warps
    __shared__ uint32_t bin_base[OSW_DIGITS];               // global position of the tile's first element of each digit
    __shared__ uint32_t wsum[NW];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, tile = blockIdx.x;
    const uint32_t n = n_ptr ? *n_ptr : n_host;
    const uint32_t start = tile * OSW_TILE;
    if (start >= n) return;
    {
        uint4 *z = reinterpret_cast<uint4 *>(&cntw[0][0]);
        for (uint32_t i = tid; i < NW * OSW_DIGITS / 8; i += OSW_THREADS) z[i] = make_uint4(0, 0, 0, 0);
    }
    // base of (tile, digit) for digits 4*tid .. 4*tid+3: rows of earlier chunks + rows of earlier tiles of this chunk
    uint32_t acc[4] = {0, 0, 0, 0};
    {
        const uint32_t chunk = tile >> chunk_shift;
        const uint4 *crow = reinterpret_cast<const uint4 *>(C) + tid;
        if (cx) { const uint4 v = crow[static_cast<size_t>(chunk) * (OSW_DIGITS / 4)]; acc[0] = v.x; acc[1] = v.y; acc[2] = v.z; acc[3] = v.w; }
        else {
#pragma unroll 8
            for (uint32_t c = 0; c < chunk; c++) { const uint4 v = crow[static_cast<size_t>(c) * (OSW_DIGITS / 4)]; acc[0] += v.x; acc[1] += v.y; acc[2] += v.z; acc[3] += v.w; }
        }
        if (H32 != nullptr) {
            const uint4 *hrow = reinterpret_cast<const uint4 *>(H32) + tid;
#pragma unroll 8
            for (uint32_t t = chunk << chunk_shift; t < tile; t++) { const uint4 v = hrow[static_cast<size_t>(t) * (OSW_DIGITS / 4)]; acc[0] += v.x; acc[1] += v.y; acc[2] += v.z; acc[3] += v.w; }
        } else {
            const ushort4 *hrow = reinterpret_cast<const ushort4 *>(H) + tid;
#pragma unroll 16
            for (uint32_t t = chunk << chunk_shift; t < tile; t++) { const ushort4 v = hrow[static_cast<size_t>(t) * (OSW_DIGITS / 4)]; acc[0] += v.x; acc[1] += v.y; acc[2] += v.z; acc[3] += v.w; }
        }
    }
    const uint4 g4 = cx ? make_uint4(0, 0, 0, 0) : reinterpret_cast<const uint4 *>(ctl_counts)[tid];
    const uint32_t gsum = g4.x + g4.y + g4.z + g4.w;
    uint32_t incl = gsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(FULL, incl, o); if (lane >= static_cast<uint32_t>(o)) incl += v; }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();

    // ---- stable in-tile ranks: warp w owns [start + w*32*ITEMS, +32*ITEMS), 32 consecutive elements per round --------
    K k[OSW_ITEMS];
    uint32_t rk[OSW_ITEMS];
#pragma unroll
    for (uint32_t r = 0; r < OSW_ITEMS; r++) {
        const uint32_t idx = start + warp * (32 * OSW_ITEMS) + r * 32 + lane;
        k[r] = idx < n ? keys_in[idx] : K(0);
    }
#pragma unroll
    for (uint32_t r = 0; r < OSW_ITEMS; r++) {
        const uint32_t idx = start + warp * (32 * OSW_ITEMS) + r * 32 + lane;
        const bool valid = idx < n && !(skip_invalid && k[r] == static_cast<K>(~K(0)));
        rk[r] = 0xffffffffu; // 0xffffffff: not an element
        if (!__any_sync(FULL, valid)) continue; // (the survivors of the streaming pass sit at the front of every 256-position tile: half of the rounds are padding)
        const uint32_t d = valid ? (static_cast<uint32_t>(k[r] >> shift) & (OSW_DIGITS - 1u)) : OSW_DIGITS + lane;
        const uint32_t mask = __match_any_sync(FULL, d);
        const uint32_t leader = static_cast<uint32_t>(__ffs(mask) - 1);
        uint32_t before = 0;
        if (valid && lane == leader) { before = cntw[warp][d]; cntw[warp][d] = static_cast<uint16_t>(before + __popc(mask)); } // one lane per digit and round
        before = __shfl_sync(FULL, before, leader);
        if (valid) rk[r] = before + __popc(mask & lanemask_lt());
        __syncwarp(); // the next round's leaders read what this round's leaders wrote
    }
    __syncthreads();
    {
        uint32_t gb = incl - gsum;
#pragma unroll
        for (uint32_t w = 0; w < NW; w++) if (w < warp) gb += wsum[w];
        uint32_t run[4] = {0, 0, 0, 0};
#pragma unroll
        for (uint32_t w = 0; w < NW; w++) { // exclusive offsets over the warps
            ushort4 *row = reinterpret_cast<ushort4 *>(&cntw[w][0]);
            const ushort4 c = row[tid];
            row[tid] = make_ushort4(static_cast<uint16_t>(run[0]), static_cast<uint16_t>(run[1]), static_cast<uint16_t>(run[2]), static_cast<uint16_t>(run[3]));
            run[0] += c.x; run[1] += c.y; run[2] += c.z; run[3] += c.w;
        }
        const uint32_t gc[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
        for (int q = 0; q < 4; q++) { bin_base[tid * 4 + q] = (region_stride ? (tid * 4 + q) * region_stride : gb) + acc[q]; gb += gc[q]; }
    }
    __syncthreads();
#pragma unroll
    for (uint32_t r = 0; r < OSW_ITEMS; r++) {
        const uint32_t idx = start + warp * (32 * OSW_ITEMS) + r * 32 + lane;
        if (rk[r] != 0xffffffffu) {
            const uint32_t d = static_cast<uint32_t>(k[r] >> shift) & (OSW_DIGITS - 1u);
            const uint32_t dst = bin_base[d] + cntw[warp][d] + rk[r];
            if (region_stride && dst - d * region_stride >= region_stride) { rk[r] = 0xffffffffu; continue; } // region overflow (the caller sees the count)
            if (keys_out != nullptr) keys_out[dst] = k[r];
            if (vals_out != nullptr) vals_out[dst] = idx;
            rk[r] = dst;
        }
    }
    if constexpr (RBYTES > 0) { // records: read in index order (coalesced), written next to the other records of their bin
        using V = typename std::conditional<RBYTES % 16 == 0, uint4, uint2>::type;
        constexpr uint32_t NV = RBYTES / sizeof(V);
#pragma unroll
        for (uint32_t r = 0; r < OSW_ITEMS; r++) {
            const uint32_t idx = start + warp * (32 * OSW_ITEMS) + r * 32 + lane;
            if (rk[r] != 0xffffffffu) {
                const V *src = reinterpret_cast<const V *>(payload_in + static_cast<size_t>(idx) * RBYTES);
                V *dstp = reinterpret_cast<V *>(payload_out + static_cast<size_t>(rk[r]) * RBYTES);
                V tmp[NV];
#pragma unroll
                for (uint32_t q = 0; q < NV; q++) tmp[q] = src[q];
#pragma unroll
                for (uint32_t q = 0; q < NV; q++) dstp[q] = tmp[q];
            }
        }
    } else if constexpr (RBYTES < 0) {
#pragma unroll 1
        for (uint32_t r = 0; r < OSW_ITEMS; r++) {
            const uint32_t idx = start + warp * (32 * OSW_ITEMS) + r * 32 + lane;
            if (rk[r] != 0xffffffffu) {
                const uint64_t *src = reinterpret_cast<const uint64_t *>(payload_in + static_cast<size_t>(idx) * payload_bytes);
                uint64_t *dstp = reinterpret_cast<uint64_t *>(payload_out + static_cast<size_t>(rk[r]) * payload_bytes);
                for (uint32_t q = 0; q < payload_bytes / 8; q++) dstp[q] = src[q];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// k_shard_scatter: the partition pass of wfb_shard_lift when there are only a few bins (destination GPUs): stable
// partition of the lifted records by dest[i] (INVALID_SLOT = dropped) into fixed-capacity regions. Same tiles and the
// same per-tile counts (H rows, chunk sums C) as k_wide_scatter, but the ranks come from one ballot per bin and round
// and the records of a bin leave a warp in runs (coalesced). nbins <= 32.
// ------------------------------------------------------------------------------------------------------
template <int RBYTES>
__global__ void __launch_bounds__(OSW_THREADS) k_shard_scatter(const uint32_t *__restrict__ dest, uint32_t n, uint32_t nbins, uint32_t chunk_shift,
                                                               const uint16_t *__restrict__ H, const uint32_t *__restrict__ C,
                                                               const unsigned char *__restrict__ payload_in, unsigned char *__restrict__ payload_out,
                                                               uint32_t region_stride)
{
    constexpr uint32_t NW = OSW_THREADS / 32;
    __shared__ uint32_t base[32];        // first output position of bin b for this tile
    __shared__ uint32_t wtot[NW][32];    // items of bin b held by warp w -> exclusive over the warps
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, tile = blockIdx.x;
    const uint32_t start = tile * OSW_TILE;
    if (start >= n) return;
    if (tid < 32) {
        uint32_t acc = 0;
        if (tid < nbins) {
            const uint32_t chunk = tile >> chunk_shift;
            for (uint32_t c = 0; c < chunk; c++) acc += C[static_cast<size_t>(c) * OSW_DIGITS + tid];
            for (uint32_t t = chunk << chunk_shift; t < tile; t++) acc += H[static_cast<size_t>(t) * OSW_DIGITS + tid];
        }
        base[tid] = tid * region_stride + acc;
    }
    // warp w owns [start + w*32*ITEMS, +32*ITEMS), 32 consecutive positions per round; lane b keeps bin b's running count
    uint32_t running = 0;
    uint16_t code[OSW_ITEMS]; // bin (5 bits) | rank inside the warp << 5; 0xffff: dropped
#pragma unroll
    for (uint32_t r = 0; r < OSW_ITEMS; r++) {
        const uint32_t idx = start + warp * (32 * OSW_ITEMS) + r * 32 + lane;
        const uint32_t d = idx < n ? dest[idx] : INVALID_SLOT;
        const uint32_t any = __ballot_sync(FULL, d < nbins);
        code[r] = 0xffffu;
        if (any == 0) continue; // (survivors are at the front of every 256-position tile: the tail rounds are empty)
        uint32_t mine = 0, add = 0;
        for (uint32_t b = 0; b < nbins; b++) {
            const uint32_t bal = __ballot_sync(FULL, d == b);
            if (d == b) mine = __popc(bal & lanemask_lt());
            if (lane == b) add = __popc(bal);
        }
        const uint32_t before = __shfl_sync(FULL, running, d < nbins ? d : 0u);
        if (d < nbins) code[r] = static_cast<uint16_t>(d | ((before + mine) << 5));
        running += add;
    }
    wtot[warp][lane] = running;
    __syncthreads();
    if (tid < 32) { // exclusive offsets over the warps
        uint32_t run = 0;
#pragma unroll
        for (uint32_t w = 0; w < NW; w++) { const uint32_t c = wtot[w][tid]; wtot[w][tid] = run; run += c; }
    }
    __syncthreads();
    using V = typename std::conditional<RBYTES % 16 == 0, uint4, uint2>::type;
    constexpr uint32_t NV = RBYTES / sizeof(V);
#pragma unroll
    for (uint32_t r = 0; r < OSW_ITEMS; r++) {
        if (code[r] != 0xffffu) {
            const uint32_t idx = start + warp * (32 * OSW_ITEMS) + r * 32 + lane;
            const uint32_t b = code[r] & 31u, rank = code[r] >> 5;
            const uint32_t local = base[b] - b * region_stride + wtot[warp][b] + rank; // position inside the region
            if (local < region_stride) {
                const V *src = reinterpret_cast<const V *>(payload_in + static_cast<size_t>(idx) * RBYTES);
                V *dstp = reinterpret_cast<V *>(payload_out + (static_cast<size_t>(b) * region_stride + local) * RBYTES);
                V tmp[NV];
#pragma unroll
                for (uint32_t q = 0; q < NV; q++) tmp[q] = src[q];
#pragma unroll
                for (uint32_t q = 0; q < NV; q++) dstp[q] = tmp[q];
            }
        }
    }
}

// multi-GPU exchange, push with SMs: every peer's slice of records (16-byte words), slots and run lengths (4-byte words) is
// stored straight into that peer's mapped receive buffer over NVLink. MG_PUSH_CTAS CTAs per peer share a slice.
constexpr uint32_t MG_PUSH_CTAS = 16, MG_PUSH_THREADS = 256; // (one CTA moves ~10 GB/s over NVLink: stores to a peer are latency-bound)
struct MgPush {
    const uint4 *rec_src[MAX_SHARDS]; uint4 *rec_dst[MAX_SHARDS]; uint32_t rec_n16[MAX_SHARDS];
    const uint32_t *slot_src[MAX_SHARDS]; uint32_t *slot_dst[MAX_SHARDS]; uint32_t slot_n[MAX_SHARDS];
    const uint32_t *bin_src[MAX_SHARDS]; uint32_t *bin_dst[MAX_SHARDS]; uint32_t bin_n;
};
static __global__ void __launch_bounds__(MG_PUSH_THREADS) k_mg_push(const __grid_constant__ MgPush a)
{
    const uint32_t peer = blockIdx.x / MG_PUSH_CTAS, part = blockIdx.x % MG_PUSH_CTAS;
    const uint32_t t = part * MG_PUSH_THREADS + threadIdx.x, stride = MG_PUSH_CTAS * MG_PUSH_THREADS;
    {
        const uint4 *src = a.rec_src[peer]; uint4 *dst = a.rec_dst[peer]; const uint32_t n = a.rec_n16[peer];
        uint32_t i = t;
        for (; i + 7 * stride < n; i += 8 * stride) { // eight independent 16-byte loads, then eight stores in flight per thread
            uint4 v[8];
#pragma unroll
            for (uint32_t q = 0; q < 8; q++) v[q] = src[i + q * stride];
#pragma unroll
            for (uint32_t q = 0; q < 8; q++) dst[i + q * stride] = v[q];
        }
        for (; i < n; i += stride) dst[i] = src[i];
    }
    {
        const uint32_t *src = a.slot_src[peer]; uint32_t *dst = a.slot_dst[peer]; const uint32_t n = a.slot_n[peer];
        for (uint32_t i = t; i < n; i += stride) dst[i] = src[i];
    }
    if (part == 0) for (uint32_t i = threadIdx.x; i < a.bin_n; i += MG_PUSH_THREADS) a.bin_dst[peer][i] = a.bin_src[peer][i];
}

// counts of the destination partition (wfb_shard_lift): counts_out[0 .. nshards) + overflow flag at [MAX_SHARDS]
static __global__ void k_shard_counts(const uint32_t *__restrict__ digit_counts, uint32_t nshards, uint32_t region_cap, uint32_t *__restrict__ counts_out)
{
    const uint32_t d = threadIdx.x;
    const uint32_t c = d < nshards ? digit_counts[d] : 0u;
    if (d < MAX_SHARDS) counts_out[d] = c;
    const uint32_t over = __ballot_sync(FULL, c > region_cap);
    if (d == 0) counts_out[MAX_SHARDS] = over ? 1u : 0u;
}


// ------------------------------------------------------------------------------------------------------
// k_ffat_update_lanes: ONE THREAD per key for the (usual) keys with few items in the segment: 65 536 keys are 65 536
// independent, short sequential folds -- enough parallelism to hide the latency of the dependent loads that bound the
// warp-per-key kernel. Each thread walks its key's items in arrival order (4 loads in flight), folds them into the open
// pane; a completed pane is written as a FlatFAT leaf and its root path recomputed in a warp-converged step (all lanes
// leave the item loop together when any of them completes a pane), fired groups go to the deferred window list.
// Keys with more than ff.light_max items are put on the heavy list for k_ffat_update (warp per key).
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t level_off(uint32_t n_leaves, uint32_t level);
__device__ __forceinline__ uint64_t batch_watermark(const uint32_t *__restrict__ batch_off, const DevBatch *__restrict__ batches,
                                                    uint32_t nbatches, uint32_t pos);
template <class P>
__device__ __forceinline__ void ffat_eval_window(const FfatDev &ff, const unsigned char *tree, uint64_t key, uint64_t gwid, uint64_t wm,
                                                 uint32_t opos, unsigned char *__restrict__ out_res, uint64_t *__restrict__ out_ts,
                                                 uint32_t out_cap, const typename P::params_t &prm);

template <class P>
__global__ void __launch_bounds__(128) k_ffat_update_lanes(const FfatDev ff, const unsigned char *__restrict__ lifted,
                                                           const uint32_t *__restrict__ sorted_pos,
                                                           const uint32_t *__restrict__ batch_off, const DevBatch *__restrict__ batches,
                                                           uint32_t nbatches, unsigned char *__restrict__ out_res,
                                                           uint64_t *__restrict__ out_ts, uint32_t out_cap, uint32_t *__restrict__ n_out,
                                                           uint32_t gather, const typename P::params_t prm)
{
    using R = typename P::result_t;
    constexpr uint32_t RB = sizeof(R);
    constexpr uint32_t U = 4; // item loads in flight per thread
    const uint32_t nslots = ff.dense ? ff.max_keys : min(*ff.n_slots, ff.max_keys);
    const uint32_t n = ff.n_leaves, logn = ff.log_leaves;
    const uint64_t P_ = ff.pane;
    const uint64_t group_items = ff.slide * ff.nb;
    const size_t tree_stride = static_cast<size_t>(2 * n - 1) * RB;

    for (uint32_t base = blockIdx.x * blockDim.x; base < nslots; base += gridDim.x * blockDim.x) { // block-uniform
        const uint32_t slot = base + threadIdx.x;
        uint32_t m = (slot < nslots) ? ff.seg_cnt[slot] : 0u;
        if (m > ff.light_max) { // heavy key: k_ffat_update takes it (its seg_cnt stays for that kernel)
            const uint32_t hi = atomicAdd(ff.n_heavy, 1u);
            if (hi < ff.max_keys) ff.heavy[hi] = slot;
            m = 0;
        }
        const bool act = m > 0;
        uint32_t off = 0; uint64_t c = 0, key = 0, g = 0, trig = 0;
        alignas(16) R acc;
        unsigned char *tree = ff.tree + static_cast<size_t>(slot) * tree_stride;
        if (act) {
            off = ff.seg_off[slot]; c = ff.cnt[slot];
            key = key_of_slot(ff, slot);
            if (c % P_ != 0) ld_rec<R>(ff.acc + static_cast<size_t>(slot) * RB, acc);
            g = (c < ff.B) ? 0 : 1 + (c - ff.B) / group_items;
            trig = ff.B + g * group_items;
        }
        uint32_t j = 0;
        while (__any_sync(FULL, act && j < m)) {
            // ---- phase 1: fold items until the open pane completes (or the key runs out of items) -------------------
            bool completed = false;
            while (act && j < m && !completed) {
                uint32_t p[U];
                alignas(16) R it[U];
                const uint32_t k = min(U, m - j);
#pragma unroll
                for (uint32_t q = 0; q < U; q++) if (q < k) p[q] = gather ? sorted_pos[off + j + q] : (off + j + q);
#pragma unroll
                for (uint32_t q = 0; q < U; q++) if (q < k) ld_rec<R>(lifted + static_cast<size_t>(p[q]) * RB, it[q]);
#pragma unroll
                for (uint32_t q = 0; q < U; q++) {
                    if (q < k && !completed) {
                        if (c % P_ == 0) acc = it[q]; else P::comb(acc, it[q], acc, prm);
                        c++; j++;
                        completed = (c % P_ == 0);
                    }
                }
            }
            // ---- phase 2 (warp-converged): new leaf, root path, fired groups -----------------------------------------------
            if (completed) {
                const uint32_t leaf = static_cast<uint32_t>((c / P_ - 1) & (n - 1));
                const uint32_t plev = ff.lazy ? 0u : logn; // (lazy: only the leaf is written)
                for (uint32_t l = 0; l < plev; l++) // siblings towards L2 first: the sequential walk below then hits
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(tree + static_cast<size_t>(level_off(n, l) + ((leaf >> l) ^ 1u)) * RB));
                alignas(16) R cur = acc;
                st_rec<R>(tree + static_cast<size_t>(leaf) * RB, cur);
                for (uint32_t l = 0; l < plev; l++) {
                    alignas(16) R s;
                    ld_rec<R>(tree + static_cast<size_t>(level_off(n, l) + ((leaf >> l) ^ 1u)) * RB, s);
                    alignas(16) R parent = cur;
                    if ((leaf >> l) & 1u) P::comb(s, cur, parent, prm); else P::comb(cur, s, parent, prm);
                    cur = parent;
                    st_rec<R>(tree + static_cast<size_t>(level_off(n, l + 1) + (leaf >> (l + 1))) * RB, cur);
                }
                if (c == trig) {
                    const uint32_t last_pos = sorted_pos[off + j - 1]; // arrival position of the triggering item
                    const uint32_t obase = atomicAdd(n_out, ff.nb);
                    bool deferred = (m - j) < ff.defer_items; // the panes this key still completes in this segment fit the spare ring leaves
                    if (deferred) {
                        const uint32_t ti = atomicAdd(ff.n_trig, 1u);
                        if (ti < ff.trig_cap) { Trigger tr; tr.key = key; tr.g = g; tr.slot = slot; tr.last_pos = last_pos; tr.obase = obase; tr.pad = 0; ff.trig[ti] = tr; }
                        else deferred = false;
                    }
                    if (!deferred) {
                        const uint64_t wm = batch_watermark(batch_off, batches, nbatches, last_pos);
                        for (uint32_t i = 0; i < ff.nb; i++)
                            ffat_eval_window<P>(ff, tree, key, g * ff.nb + i, wm, obase + i, out_res, out_ts, out_cap, prm);
                    }
                    g++; trig += group_items;
                }
            }
        }
        if (act) {
            ff.cnt[slot] = c;
            if (c % P_ != 0) st_rec<R>(ff.acc + static_cast<size_t>(slot) * RB, acc);
            ff.seg_cnt[slot] = 0;
            ff.seg_off[slot] = 0xffffffffu; // the next segment's sort records the key's first position with atomicMin
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// k_ffat_update_buckets: the window update after ONE wide partition pass (k_wide_scatter) on the top 10 bits of the
// slot. The pass leaves the segment's (slot, arrival position) pairs in 1024 buckets of at most BK_KEYS consecutive
// slots, arrival order inside a bucket. One CTA per bucket, in chunks of at most BK_CAP items (arrival order is chunk
// order). Every phase is built to need as few dependent memory round trips as possible -- the kernel is bound by
// latency, not by bytes:
//   1. every thread takes BK_IT consecutive items of the chunk; stable split by key through per-thread private key
//      counts in shared memory (count, exclusive scan over the threads, place) -> per-key runs of record indices,
//   2. the FlatFAT siblings of the leaf each key completes first in this chunk are copied to shared memory with cp.async
//      while
//   3. ONE THREAD per segment (the items of a run that fall into one pane) folds its items in arrival order straight
//      from the lifted array, eight masked loads per round trip; then one thread per key walks its segments in order:
//      completed panes become FlatFAT leaves, their root paths are recomputed (staged siblings for the first) and fired
//      groups go to the deferred window list; the last partial segment is the new open pane,
//   4. a chunk with more segments than threads (tiny panes) is folded one warp per key instead (ordered shuffle-tree
//      fold, 32 records per load).
// Per-key bookkeeping (count, position in the open pane, next leaf, next trigger, open-pane accumulator) is computed
// once per CTA by one thread per key and lives in shared memory across the chunks.
// `moved` = 1: the partition pass also moved the records (bucket b's records are lifted[boff[b] ..)); 0: records are
// gathered through the arrival positions.
// ------------------------------------------------------------------------------------------------------
#ifdef WFB_BK_TRACE
__device__ unsigned long long g_bk_trace[1024 * 8];
#define BK_MARK(i) do { if (threadIdx.x == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); g_bk_trace[blockIdx.x * 8 + (i)] = t_; } } while (0)
#else
#define BK_MARK(i) do { } while (0)
#endif
constexpr uint32_t BK_KEYS = 64;      // keys per bucket (at most)
constexpr uint32_t BK_THREADS = 128;
constexpr uint32_t BK_IT = 18;        // consecutive items per thread and chunk
constexpr uint32_t BK_CAP = BK_THREADS * BK_IT; // items per chunk
#ifndef WFB_BK_U
#define WFB_BK_U 4
#endif
#ifndef WFB_BK_MINBLOCKS
#define WFB_BK_MINBLOCKS 4
#endif
#ifndef WFB_BK_MINBLOCKS_LAZY
#define WFB_BK_MINBLOCKS_LAZY 5   // lazy FlatFAT levels: no sibling staging (32 KB of shared memory instead of 48), no path code
#endif
constexpr uint32_t BK_U = WFB_BK_U;   // record loads in flight per thread

template <class P, bool LAZY>
__global__ void __launch_bounds__(BK_THREADS, LAZY ? WFB_BK_MINBLOCKS_LAZY : WFB_BK_MINBLOCKS) k_ffat_update_buckets(const FfatDev ff, const unsigned char *__restrict__ lifted,
                                                                       const uint32_t *__restrict__ bk_slots, const uint32_t *__restrict__ bk_pos,
                                                                       const uint32_t *__restrict__ digit_counts, uint32_t shift, uint32_t moved,
                                                                       const uint32_t *__restrict__ batch_off, const DevBatch *__restrict__ batches,
                                                                       uint32_t nbatches, unsigned char *__restrict__ out_res,
                                                                       uint64_t *__restrict__ out_ts, uint32_t out_cap, uint32_t *__restrict__ n_out,
                                                                       const typename P::params_t prm)
{
    using R = typename P::result_t;
    constexpr uint32_t RB = sizeof(R);
    constexpr uint32_t NW = BK_THREADS / 32;
    constexpr uint32_t DPT = OSW_DIGITS / BK_THREADS;                  // digit counts per thread
    constexpr uint32_t HS = BK_THREADS + 2;                            // row stride of the private counts (bank-conflict padding)
    constexpr uint32_t CPB = (RB % 16 == 0) ? 16 : 8;                  // cp.async granule of a record
    constexpr uint32_t BK_SIBL = RB <= 32 ? 8 : (RB <= 48 ? 4 : (RB <= 128 ? 2 : 1)); // FlatFAT levels whose siblings are staged in shared memory (static shared memory <= 48 KB)
    static_assert(DPT % 4 == 0 && BK_KEYS == 64 && BK_THREADS == 128 && RB % 8 == 0, "layout");
    __shared__ uint32_t s_idx[BK_CAP];                 // record index of the items (into `lifted`), key-major
    __shared__ __align__(16) uint16_t hist[BK_KEYS][HS]; // items of key k among thread t's items -> exclusive over the threads;
                                                       // after the split: segment descriptors and segment results (fold phase)
    __shared__ uint32_t htot[2][BK_KEYS];              // per key: items held by threads 0..63 / 64..127
    __shared__ uint32_t kcnt[BK_KEYS], koff[BK_KEYS];  // items / first index of key k in this chunk
    __shared__ uint32_t kleft[BK_KEYS];                // items of key k still to come in this segment
    __shared__ uint32_t kcp[BK_KEYS], kleaf[BK_KEYS];  // items in the open pane, leaf the open pane will be written to
    __shared__ uint64_t kc[BK_KEYS], kg[BK_KEYS], ktt[BK_KEYS]; // count, groups fired, items until the next trigger
    __shared__ __align__(16) unsigned char kacc[BK_KEYS * RB];   // open-pane accumulator of key k
    __shared__ __align__(16) unsigned char s_sib[LAZY ? 16 : BK_KEYS * BK_SIBL * RB]; // siblings of the leaf key k completes in this chunk (lazy levels: none)
    __shared__ uint32_t s_heavy[BK_KEYS];              // keys folded by a warp in this chunk
    __shared__ uint32_t ksegb[BK_KEYS];                // first segment of key k (a segment = the items of a run that fall into one pane)
    __shared__ uint32_t misc[NW], s_boff[2], s_nheavy, s_nseg;
    constexpr bool SEG_OK = BK_THREADS * (RB + 4) <= sizeof(uint16_t) * BK_KEYS * HS; // segment results + descriptors alias the private counts
    unsigned char *seg_res = reinterpret_cast<unsigned char *>(&hist[0][0]);                       // BK_THREADS x result_t
    uint32_t *seg_desc = reinterpret_cast<uint32_t *>(seg_res + BK_THREADS * RB);                  // key | index of the segment in its run << 8

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t bucket = blockIdx.x;
    const uint32_t kpc = min(BK_KEYS, 1u << shift);    // keys of this bucket
    const uint32_t key_lo = bucket << shift;
    const uint32_t n = ff.n_leaves, logn = ff.log_leaves;
    const uint32_t P32 = static_cast<uint32_t>(ff.pane);
    const uint64_t group_items = ff.slide * ff.nb;
    const size_t tree_stride = static_cast<size_t>(2 * n - 1) * RB;

    BK_MARK(0);
    // ---- per-key bookkeeping, one thread per key (loads first: they overlap the histogram scan below) ----------------------
    uint32_t my_total = 0;
    uint64_t st_c = 0;
    alignas(16) R st_acc;
    const bool has_key = tid < kpc && key_lo + tid < ff.max_keys;
    if (has_key) {
        const uint32_t slot = key_lo + tid;
        st_c = ff.cnt[slot];
        ld_rec<R>(ff.acc + static_cast<size_t>(slot) * RB, st_acc);
    }
    if (tid < BK_KEYS) kleft[tid] = 0;
    // ---- bucket range = exclusive scan of the pass histogram --------------------------------------------------------------
    {
        uint32_t cc[DPT];
#pragma unroll
        for (uint32_t q = 0; q < DPT / 4; q++) {
            const uint4 v = reinterpret_cast<const uint4 *>(digit_counts)[tid * (DPT / 4) + q];
            cc[4 * q] = v.x; cc[4 * q + 1] = v.y; cc[4 * q + 2] = v.z; cc[4 * q + 3] = v.w;
        }
        uint32_t sum = 0;
#pragma unroll
        for (uint32_t q = 0; q < DPT; q++) sum += cc[q];
        uint32_t incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(FULL, incl, o); if (lane >= static_cast<uint32_t>(o)) incl += v; }
        if (lane == 31) misc[warp] = incl;
        __syncthreads();
        if (tid == bucket / DPT) {
            uint32_t base = incl - sum;
            for (uint32_t w = 0; w < warp; w++) base += misc[w];
            uint32_t own = 0;
#pragma unroll
            for (uint32_t q = 0; q < DPT; q++) { if (q < bucket % DPT) base += cc[q]; if (q == bucket % DPT) own = cc[q]; }
            s_boff[0] = base; s_boff[1] = base + own;
        }
        __syncthreads();
    }
    // items of every key in this stream segment (the whole bucket, all chunks): the deferral of fired groups needs them. Counting
    // them here instead of one global RED per survivor in the streaming pass is what that pass is sensitive to (+17 us per RED).
    for (uint32_t i = s_boff[0] + tid; i < s_boff[1]; i += BK_THREADS) {
        const uint32_t lk = bk_slots[i] - key_lo;
        if (lk < kpc) atomicAdd(&kleft[lk], 1u);
    }
    __syncthreads();
    if (tid < BK_KEYS) {
        const uint32_t m = has_key ? kleft[tid] : 0u;
        my_total = m;
        const uint64_t c = st_c;
        uint64_t g = 0, tt = 0; uint32_t cp = 0, leaf = 0;
        if (m) {
            cp = static_cast<uint32_t>(c % P32); leaf = static_cast<uint32_t>((c / P32) & (n - 1));
            if (c < ff.B) { g = 0; tt = ff.B - c; }
            else { g = 1 + (c - ff.B) / group_items; tt = ff.B + g * group_items - c; }
            if (cp) st_rec<R>(kacc + tid * RB, st_acc);
        }
        kleft[tid] = m; kc[tid] = c; kg[tid] = g; ktt[tid] = tt; kcp[tid] = cp; kleaf[tid] = leaf;
    }
    const uint32_t any_items = __syncthreads_or(my_total != 0);
    BK_MARK(1);
    if (!any_items) return;
    uint32_t cursor = s_boff[0];
    const uint32_t bend = s_boff[1];
    if (moved)
        for (size_t o = static_cast<size_t>(tid) * 128; o < static_cast<size_t>(bend - cursor) * RB; o += BK_THREADS * 128) // the bucket's block -> L2
            asm volatile("prefetch.global.L2 [%0];" ::"l"(lifted + static_cast<size_t>(cursor) * RB + o));

    while (cursor < bend) {
        const uint32_t nsel = min(BK_CAP, bend - cursor);
        // ---- 1. stable split of the chunk's items by key: thread t owns items [t*BK_IT, +BK_IT) --------------------------------------
        uint32_t ek[BK_IT], ep[BK_IT]; // local key (BK_KEYS = none), record index
#pragma unroll
        for (uint32_t r = 0; r < BK_IT; r++) {
            const uint32_t i = tid * BK_IT + r;
            ek[r] = BK_KEYS; ep[r] = cursor + i;
            if (i < nsel) {
                const uint32_t lk = bk_slots[cursor + i] - key_lo; // slots outside the bucket's keys (invalid slots) are dropped
                if (!moved) ep[r] = bk_pos[cursor + i];             // records still in arrival order: the index is the position
                if (lk < kpc) ek[r] = lk;
            }
        }
        {
            uint32_t *z = reinterpret_cast<uint32_t *>(&hist[0][0]);
            for (uint32_t i = tid; i < BK_KEYS * HS / 2; i += BK_THREADS) z[i] = 0;
        }
        __syncthreads();
#pragma unroll
        for (uint32_t r = 0; r < BK_IT; r++) if (ek[r] < BK_KEYS) hist[ek[r]][tid]++; // column tid is private to this thread
        __syncthreads();
        { // key (tid & 63), threads [64*(tid >> 6), +64): exclusive scan of the private counts over the threads
            const uint32_t k = tid & 63u, half = tid >> 6;
            uint16_t *row = &hist[k][half * 64];
            uint32_t run = 0;
#pragma unroll 16
            for (uint32_t i = 0; i < 64; i++) { const uint32_t c = row[i]; row[i] = static_cast<uint16_t>(run); run += c; }
            htot[half][k] = run;
        }
        __syncthreads();
        BK_MARK(2);
        if (warp == 0) { // keys lane and lane+32: chunk totals, exclusive scan over the keys, keys with long runs
            const uint32_t a0 = htot[0][lane] + htot[1][lane], a1 = htot[0][lane + 32] + htot[1][lane + 32];
            uint32_t i0 = a0, i1 = a1;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v0 = __shfl_up_sync(FULL, i0, o), v1 = __shfl_up_sync(FULL, i1, o);
                if (lane >= static_cast<uint32_t>(o)) { i0 += v0; i1 += v1; }
            }
            const uint32_t t0 = __shfl_sync(FULL, i0, 31);
            kcnt[lane] = a0; kcnt[lane + 32] = a1;
            koff[lane] = i0 - a0; koff[lane + 32] = t0 + i1 - a1;
            // segments of the runs: the items that complete the open pane, then one segment per further pane
            const uint32_t c0 = kcp[lane], c1 = kcp[lane + 32];
            const uint32_t f0 = min(a0, P32 - c0), f1 = min(a1, P32 - c1);
            const uint32_t n0 = a0 ? 1u + (a0 - f0 + P32 - 1) / P32 : 0u, n1 = a1 ? 1u + (a1 - f1 + P32 - 1) / P32 : 0u;
            uint32_t s0 = n0, s1 = n1;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v0 = __shfl_up_sync(FULL, s0, o), v1 = __shfl_up_sync(FULL, s1, o);
                if (lane >= static_cast<uint32_t>(o)) { s0 += v0; s1 += v1; }
            }
            const uint32_t st0 = __shfl_sync(FULL, s0, 31), nseg = st0 + __shfl_sync(FULL, s1, 31);
            ksegb[lane] = s0 - n0; ksegb[lane + 32] = st0 + s1 - n1;
            // more segments than threads (tiny panes): every run of the chunk is folded by a warp instead
            const bool fallback = !SEG_OK || nseg > BK_THREADS;
            const bool h0 = fallback && a0 != 0, h1 = fallback && a1 != 0;
            const uint32_t b0 = __ballot_sync(FULL, h0), b1 = __ballot_sync(FULL, h1);
            if (h0) s_heavy[__popc(b0 & lanemask_lt())] = lane;
            if (h1) s_heavy[__popc(b0) + __popc(b1 & lanemask_lt())] = lane + 32;
            if (lane == 0) { s_nheavy = __popc(b0) + __popc(b1); s_nseg = fallback ? 0u : nseg; }
        }
        __syncthreads();
        {
            const uint32_t hb = tid >> 6;
#pragma unroll
            for (uint32_t r = 0; r < BK_IT; r++) {
                const uint32_t k = ek[r];
                if (k < BK_KEYS) {
                    const uint32_t rank = hist[k][tid];
                    hist[k][tid] = static_cast<uint16_t>(rank + 1);
                    s_idx[koff[k] + (hb ? htot[0][k] : 0u) + rank] = ep[r];
                }
            }
        }
        __syncthreads(); // the private counts are dead: their memory now holds the segment descriptors / results
        // ---- 2. segment descriptors; siblings of the first leaf each key completes -> shared memory (asynchronous) -----------------
        const uint32_t nseg = s_nseg;
        bool sib_staged = false;
        if (tid < BK_KEYS && nseg != 0) {
            const uint32_t m = kcnt[tid];
            if (m != 0) {
                const uint32_t cp0 = kcp[tid], first = min(m, P32 - cp0), ns = 1u + (m - first + P32 - 1) / P32, sb = ksegb[tid];
                for (uint32_t j = 0; j < ns; j++) seg_desc[sb + j] = tid | (j << 8);
                if (!LAZY && cp0 + first == P32) {
                    const uint32_t leaf = kleaf[tid];
                    const unsigned char *tr = ff.tree + static_cast<size_t>(key_lo + tid) * tree_stride;
                    for (uint32_t l = 0; l < min(logn, BK_SIBL); l++) {
                        const unsigned char *src = tr + static_cast<size_t>(level_off(n, l) + ((leaf >> l) ^ 1u)) * RB;
                        unsigned char *dst = s_sib + (tid * BK_SIBL + l) * RB;
#pragma unroll
                        for (uint32_t q = 0; q < RB / CPB; q++) cp_async<CPB>(dst + q * CPB, src + q * CPB);
                    }
                    sib_staged = true;
                }
            }
        }
        __syncthreads();
        BK_MARK(3);
        // ---- 3a. one thread per segment: plain ordered fold of at most one pane of items, BK_U masked loads per round trip ---------
        if (tid < nseg) {
            const uint32_t d = seg_desc[tid], k = d & 255u, j = d >> 8;
            const uint32_t m = kcnt[k], cp0 = kcp[k], first = min(m, P32 - cp0);
            const uint32_t start = j == 0 ? 0u : first + (j - 1) * P32;
            uint32_t seg = j == 0 ? first : min(P32, m - start);
            bool fresh = j != 0 || cp0 == 0;
            const uint32_t *ip = s_idx + koff[k] + start;
            alignas(16) R acc;
            if (!fresh) ld_rec<R>(kacc + k * RB, acc);
            while (seg != 0) {
                alignas(16) R it[BK_U];
                const uint32_t kk = min(seg, BK_U);
#pragma unroll
                for (uint32_t q = 0; q < BK_U; q++) if (q < kk) ld_rec<R>(lifted + static_cast<size_t>(ip[q]) * RB, it[q]);
                if (fresh) acc = it[0]; else P::comb(acc, it[0], acc, prm);
                fresh = false;
#pragma unroll
                for (uint32_t q = 1; q < BK_U; q++) if (q < kk) P::comb(acc, it[q], acc, prm);
                seg -= kk; ip += kk;
            }
            st_rec<R>(seg_res + tid * RB, acc);
        }
        __syncthreads();
        // ---- 3b. one thread per key, its segments in order: completed panes -> leaf + root path + fired group; the rest is the open
        // pane. The loop over the segments is warp-uniform: a group that cannot be deferred (another pane of the key completes
        // in this stream segment and would overwrite ring leaves the windows still need) is evaluated by the whole warp at once.
        if (tid < BK_KEYS && nseg != 0) { // warps 0 and 1, whole
            const uint32_t k = tid, slot = key_lo + k, m = kcnt[k];
            const uint32_t cp0 = kcp[k], first = min(m, P32 - cp0), ns = m ? 1u + (m - first + P32 - 1) / P32 : 0u, sb = ksegb[k];
            uint32_t leafi = kleaf[k], new_cp = cp0, consumed = 0;
            uint64_t g = kg[k], tt = ktt[k]; // tt: index (1-based) of the run's item that fires the next group
            const uint32_t left0 = kleft[k];
            unsigned char *tree = ff.tree + static_cast<size_t>(slot) * tree_stride;
            const uint64_t key = key_of_slot(ff, slot < ff.max_keys ? slot : 0u);
            for (uint32_t j = 0; __any_sync(FULL, j < ns); j++) {
                bool eval_now = false;
                uint32_t ev_obase = 0, ev_pos = 0;
                uint64_t ev_g = 0;
                if (j < ns) {
                    const uint32_t len = j == 0 ? first : min(P32, m - consumed);
                    consumed += len;
                    alignas(16) R cur;
                    ld_rec<R>(seg_res + (sb + j) * RB, cur);
                    const bool completes = (j == 0 ? cp0 + len : len) == P32;
                    if (!completes) { st_rec<R>(kacc + k * RB, cur); new_cp = (j == 0 ? cp0 : 0u) + len; } // (only the last segment)
                    else {
                        new_cp = 0;
                        const uint32_t leaf = leafi;
                        leafi = (leafi + 1) & (n - 1);
                        st_rec<R>(tree + static_cast<size_t>(leaf) * RB, cur);
                        if (sib_staged) cp_async_wait_all();
                        for (uint32_t l0 = 0; l0 < (LAZY ? 0u : logn); l0 += 4) { // siblings of four levels per round trip (none of them is on the path)
                            alignas(16) R sbl[4];
#pragma unroll
                            for (uint32_t q = 0; q < 4; q++) {
                                const uint32_t l = l0 + q;
                                if (l < logn) {
                                    if (sib_staged && l < BK_SIBL) ld_rec<R>(s_sib + (k * BK_SIBL + l) * RB, sbl[q]);
                                    else ld_rec<R>(tree + static_cast<size_t>(level_off(n, l) + ((leaf >> l) ^ 1u)) * RB, sbl[q]);
                                }
                            }
#pragma unroll
                            for (uint32_t q = 0; q < 4; q++) {
                                const uint32_t l = l0 + q;
                                if (l < logn) {
                                    alignas(16) R parent = cur;
                                    if ((leaf >> l) & 1u) P::comb(sbl[q], cur, parent, prm); else P::comb(cur, sbl[q], parent, prm);
                                    cur = parent;
                                    st_rec<R>(tree + static_cast<size_t>(level_off(n, l + 1) + (leaf >> (l + 1))) * RB, cur);
                                }
                            }
                        }
                        sib_staged = false; // the next pane of the same run reads the tree it has just written
                        if (consumed == tt) {
                            const uint32_t lp = s_idx[koff[k] + consumed - 1];
                            const uint32_t last_pos = moved ? bk_pos[lp] : lp; // arrival position of the triggering item
                            const uint32_t obase = atomicAdd(n_out, ff.nb);
                            bool deferred = (left0 - consumed) < ff.defer_items; // the panes this key still completes in this segment fit the spare ring leaves
                            if (deferred) {
                                const uint32_t ti = atomicAdd(ff.n_trig, 1u);
                                if (ti < ff.trig_cap) { Trigger tr; tr.key = key; tr.g = g; tr.slot = slot; tr.last_pos = last_pos; tr.obase = obase; tr.pad = 0; ff.trig[ti] = tr; }
                                else deferred = false;
                            }
                            if (!deferred) { eval_now = true; ev_obase = obase; ev_pos = last_pos; ev_g = g; }
                            g++; tt += group_items;
                        }
                    }
                }
                // groups to evaluate before their key's next pane: one after the other, the warp's lanes share the windows
                uint32_t pend = __ballot_sync(FULL, eval_now);
                while (pend) {
                    const int src = __ffs(pend) - 1;
                    pend &= pend - 1;
                    const uint32_t e_slot = __shfl_sync(FULL, slot, src), e_obase = __shfl_sync(FULL, ev_obase, src), e_pos = __shfl_sync(FULL, ev_pos, src);
                    const uint64_t e_key = __shfl_sync(FULL, key, src), e_g = __shfl_sync(FULL, ev_g, src);
                    const unsigned char *e_tree = ff.tree + static_cast<size_t>(e_slot) * tree_stride;
                    const uint64_t wm = batch_watermark(batch_off, batches, nbatches, e_pos);
                    for (uint32_t i = lane; i < ff.nb; i += 32)
                        ffat_eval_window<P>(ff, e_tree, e_key, e_g * ff.nb + i, wm, e_obase + i, out_res, out_ts, out_cap, prm);
                }
                __syncwarp(); // the evaluated windows read tree nodes another lane has just written, and it may overwrite them next
            }
            if (sib_staged) cp_async_wait_all(); // (a staged key always completes its first segment: nothing is pending here)
            if (m != 0) { kc[k] += m; kg[k] = g; ktt[k] = tt - m; kcp[k] = new_cp; kleaf[k] = leafi; kleft[k] = left0 - m; }
        }
        BK_MARK(4);
        // ---- 3. long runs: one warp per key, ordered shuffle-tree fold of 32 records per load ----------------------------------------
        if (s_nheavy != 0) { // (block-uniform: written before the last barrier)
            __syncthreads();
            for (uint32_t h = warp; h < s_nheavy; h += NW) {
                const uint32_t k = s_heavy[h];
                const uint32_t m = kcnt[k], off = koff[k], slot = key_lo + k;
                uint64_t c = kc[k], g = kg[k], tt = ktt[k];
                uint32_t cp = kcp[k], leafi = kleaf[k], left = kleft[k];
                const uint64_t key = key_of_slot(ff, slot);
                unsigned char *tree = ff.tree + static_cast<size_t>(slot) * tree_stride;
                alignas(16) R acc;
                if (cp) ld_rec<R>(kacc + k * RB, acc);
                alignas(16) R cur_rec;
                if (lane < m) ld_rec<R>(lifted + static_cast<size_t>(s_idx[off + lane]) * RB, cur_rec);
                uint32_t j = 0;
                while (j < m) {
                    const uint32_t cnt = min(32u, m - j);
                    alignas(16) R nxt_rec;
                    if (j + 32 + lane < m) ld_rec<R>(lifted + static_cast<size_t>(s_idx[off + j + 32 + lane]) * RB, nxt_rec);
                    uint32_t lo = 0;
                    while (lo < cnt) { // sub-ranges of the 32 records that fall into one pane
                        const uint32_t hi = min(cnt, lo + (P32 - cp));
                        alignas(16) R r = cur_rec;
#pragma unroll
                        for (uint32_t o = 1; o < 32; o <<= 1) {
                            const R other = shfl_down_rec<R>(r, o);
                            if (lane >= lo && lane + o < hi) P::comb(r, other, r, prm);
                        }
                        r = shfl_rec<R>(r, lo);
                        if (cp == 0) acc = r; else P::comb(acc, r, acc, prm);
                        const uint32_t take = hi - lo;
                        cp += take; c += take; left -= take; tt -= take;
                        if (cp == P32) { // pane complete -> leaf + root path
                            cp = 0;
                            const uint32_t leaf = leafi;
                            leafi = (leafi + 1) & (n - 1);
                            alignas(16) R sib;
                            const uint32_t plev = LAZY ? 0u : logn; // (lazy: only the leaf is written)
                            if (lane < plev) ld_rec<R>(tree + static_cast<size_t>(level_off(n, lane) + ((leaf >> lane) ^ 1u)) * RB, sib);
                            alignas(16) R cur = acc;
                            if (lane == 0) st_rec<R>(tree + static_cast<size_t>(leaf) * RB, cur);
                            for (uint32_t l = 0; l < plev; l++) {
                                const R sb = shfl_rec<R>(sib, l);
                                alignas(16) R parent = cur;
                                if ((leaf >> l) & 1u) P::comb(sb, cur, parent, prm); else P::comb(cur, sb, parent, prm);
                                cur = parent;
                                if (lane == 0) st_rec<R>(tree + static_cast<size_t>(level_off(n, l + 1) + (leaf >> (l + 1))) * RB, cur);
                            }
                            __syncwarp();
                            if (tt == 0) {
                                const uint32_t last_pos = moved ? bk_pos[s_idx[off + j + hi - 1]] : s_idx[off + j + hi - 1]; // arrival position of the triggering item
                                uint32_t obase = 0;
                                if (lane == 0) obase = atomicAdd(n_out, ff.nb);
                                obase = __shfl_sync(FULL, obase, 0);
                                bool deferred = left < ff.defer_items; // the panes this key still completes in this segment fit the spare ring leaves
                                if (deferred) {
                                    uint32_t ti = 0;
                                    if (lane == 0) ti = atomicAdd(ff.n_trig, 1u);
                                    ti = __shfl_sync(FULL, ti, 0);
                                    if (ti < ff.trig_cap) {
                                        if (lane == 0) { Trigger tr; tr.key = key; tr.g = g; tr.slot = slot; tr.last_pos = last_pos; tr.obase = obase; tr.pad = 0; ff.trig[ti] = tr; }
                                    } else deferred = false;
                                }
                                if (!deferred) {
                                    const uint64_t wm = batch_watermark(batch_off, batches, nbatches, last_pos);
                                    for (uint32_t i = lane; i < ff.nb; i += 32)
                                        ffat_eval_window<P>(ff, tree, key, g * ff.nb + i, wm, obase + i, out_res, out_ts, out_cap, prm);
                                }
                                g++; tt = group_items;
                                __syncwarp();
                            }
                        }
                        lo = hi;
                    }
                    j += cnt;
                    cur_rec = nxt_rec;
                }
                if (lane == 0) {
                    kc[k] = c; kg[k] = g; ktt[k] = tt; kcp[k] = cp; kleaf[k] = leafi; kleft[k] = left;
                    if (cp) st_rec<R>(kacc + k * RB, acc);
                }
            }
        }
        cursor += nsel;
        __syncthreads(); // the next chunk overwrites the shared buffers
    }
    BK_MARK(5);
    // ---- keys' state back as contiguous blocks -----------------------------------------------------------------------------------
    if (tid < kpc && my_total) {
        const uint32_t slot = key_lo + tid;
        ff.cnt[slot] = kc[tid];
        if (kcp[tid]) { alignas(16) R a; ld_rec<R>(kacc + tid * RB, a); st_rec<R>(ff.acc + static_cast<size_t>(slot) * RB, a); }
    }
    BK_MARK(6);
}

// ------------------------------------------------------------------------------------------------------
// k_ffat_update_stream: the window update as 32 independent STREAMS per warp -- no in-bucket sort, no per-key runs, no block
// barriers in the loop. One CTA (2 warps) per bucket of the wide partition; warp w owns the bucket's keys [32 w, 32 w + 32) and
// LANE k OF THE WARP IS KEY 32 w + k: the key's count, open pane (accumulator in registers), next leaf and next trigger live in
// that lane's registers for the whole kernel, and the lane folds the key's items itself, in arrival order:
//   producer  every warp reads the bucket's (slot, position) pairs in arrival order (128 per step, coalesced, prefetched one step
//             ahead; the two warps share the lines through L1). The positions of the warp's own items go to PER-KEY FIFO queues in
//             shared memory (match_any gives an item its rank among the group's items of the same key, the key lane's tail comes
//             by shuffle): arrival order per key is queue order.
//   consumer  in a round every lane pops the next position of ITS key, starts the copy of that record global -> shared (cp.async
//             into a private 3-deep ring: no registers, no other lane involved) and folds the record it asked for three rounds ago
//             into the open pane. All 32 lanes work on 32 different keys: a round costs ~40 warp instructions for up to 32
//             items, and three gathers per lane are in flight. Rounds run whenever the queues hold two items per key on average.
//   a lane that completes a pane writes the FlatFAT leaf and recomputes the root path -- with the siblings of the first leaf each
//   key completes staged in shared memory at kernel start (cp.async by the key's own lane), so the usual completion issues no
//   dependent global load; a fired group is deferred to k_ffat_windows; should the same key complete ANOTHER pane later in this
//   call (it would overwrite ring leaves the deferred windows still read), the pending group is evaluated first by the whole warp
//   and its list entry voided. No per-key item counts of the segment are needed for that decision (the bucket kernel counts first).
// Built for panes of at least a few items (a pane per item would make every round a path update): the host selects
// k_ffat_update_buckets otherwise.
// ------------------------------------------------------------------------------------------------------
#ifndef WFB_ST_MINBLOCKS
#define WFB_ST_MINBLOCKS 7
#endif
constexpr uint32_t ST_THREADS = 64, ST_WARPS = ST_THREADS / 32;
constexpr uint32_t ST_QCAP = 32;   // positions a key's queue holds (a whole group of 32 items of one key fits an empty queue)
constexpr uint32_t ST_FLY = 3;     // gathers in flight per lane
constexpr uint32_t ST_BACKLOG = 64; // consumer rounds run while the warp's queues hold at least this many items (two per key: few idle lanes)
constexpr uint32_t ST_NONE = 0xffffffffu;
static_assert(ST_WARPS * 32 == BK_KEYS, "one key per lane");

template <class P>
__global__ void __launch_bounds__(ST_THREADS, WFB_ST_MINBLOCKS) k_ffat_update_stream(const FfatDev ff, const unsigned char *__restrict__ lifted,
                                                                     const uint32_t *__restrict__ bk_slots, const uint32_t *__restrict__ bk_pos,
                                                                     const uint32_t *__restrict__ digit_counts, uint32_t shift,
                                                                     const uint32_t *__restrict__ batch_off, const DevBatch *__restrict__ batches,
                                                                     uint32_t nbatches, unsigned char *__restrict__ out_res,
                                                                     uint64_t *__restrict__ out_ts, uint32_t out_cap, uint32_t *__restrict__ n_out,
                                                                     const typename P::params_t prm)
{
    using R = typename P::result_t;
    constexpr uint32_t RB = sizeof(R);
    constexpr uint32_t DPT = OSW_DIGITS / ST_THREADS;
    constexpr uint32_t CPB = (RB % 16 == 0) ? 16 : 8;
    constexpr uint32_t SIBL = RB <= 32 ? 8 : (RB <= 64 ? 4 : (RB <= 128 ? 2 : 1)); // FlatFAT levels whose siblings are staged
    static_assert(DPT % 4 == 0 && RB % 8 == 0, "layout");
    __shared__ __align__(16) unsigned char s_sib[BK_KEYS * SIBL * RB];          // siblings of the leaf key k completes first (private to the key's lane)
    __shared__ __align__(16) unsigned char s_stage[ST_THREADS * ST_FLY * RB];   // records in flight: ST_FLY per lane (private to the lane)
    __shared__ uint32_t q_pos[ST_THREADS][ST_QCAP + 1];                         // per-key FIFO of arrival positions (+1: bank-conflict padding)
    __shared__ uint32_t misc[ST_WARPS], s_boff[2];

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t bucket = blockIdx.x;
    const uint32_t kpc = min(BK_KEYS, 1u << shift);
    const uint32_t key_lo = bucket << shift;
    const uint32_t n = ff.n_leaves, logn = ff.log_leaves;
    const uint32_t P32 = static_cast<uint32_t>(ff.pane);
    const uint64_t group_items = ff.slide * ff.nb;
    const size_t tree_stride = static_cast<size_t>(2 * n - 1) * RB;

    // ---- my key's state (loads first: they overlap the scan below) ------------------------------------------------------------------
    const uint32_t my_k = tid; // local key of this lane
    const bool has_key = my_k < kpc && key_lo + my_k < ff.max_keys;
    const uint32_t my_slot = key_lo + my_k;
    unsigned char *const my_tree = ff.tree + static_cast<size_t>(has_key ? my_slot : 0u) * tree_stride;
    uint64_t st_c = 0;
    alignas(16) R acc;
    if (has_key) {
        st_c = ff.cnt[my_slot];
        ld_rec<R>(ff.acc + static_cast<size_t>(my_slot) * RB, acc);
    }
    // ---- bucket range = exclusive scan of the pass histogram ------------------------------------------------------------------------
    {
        uint32_t cc[DPT];
#pragma unroll
        for (uint32_t q = 0; q < DPT / 4; q++) {
            const uint4 v = reinterpret_cast<const uint4 *>(digit_counts)[tid * (DPT / 4) + q];
            cc[4 * q] = v.x; cc[4 * q + 1] = v.y; cc[4 * q + 2] = v.z; cc[4 * q + 3] = v.w;
        }
        uint32_t sum = 0;
#pragma unroll
        for (uint32_t q = 0; q < DPT; q++) sum += cc[q];
        uint32_t incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(FULL, incl, o); if (lane >= static_cast<uint32_t>(o)) incl += v; }
        if (lane == 31) misc[warp] = incl;
        __syncthreads();
        if (tid == bucket / DPT) {
            uint32_t base = incl - sum;
            for (uint32_t w = 0; w < warp; w++) base += misc[w];
            uint32_t own = 0;
#pragma unroll
            for (uint32_t q = 0; q < DPT; q++) { if (q < bucket % DPT) base += cc[q]; if (q == bucket % DPT) own = cc[q]; }
            s_boff[0] = base; s_boff[1] = base + own;
        }
    }
    uint64_t g = 0, tt = 0;          // groups fired | ordinal (1-based, among the key's items of this call) of the item that fires the next group
    uint32_t cp = 0, leaf = 0, cons = 0, pend = ST_NONE; // items in the open pane | leaf it becomes | items consumed in this call | deferred group of this call
    bool staged = false;
    if (has_key) {
        const uint64_t c = st_c;
        cp = static_cast<uint32_t>(c % P32); leaf = static_cast<uint32_t>((c / P32) & (n - 1));
        if (c < ff.B) { g = 0; tt = ff.B - c; }
        else { g = 1 + (c - ff.B) / group_items; tt = ff.B + g * group_items - c; }
        for (uint32_t l = 0; l < (ff.lazy ? 0u : min(logn, SIBL)); l++) {
            const unsigned char *src = my_tree + static_cast<size_t>(level_off(n, l) + ((leaf >> l) ^ 1u)) * RB;
            unsigned char *dst = s_sib + (my_k * SIBL + l) * RB;
#pragma unroll
            for (uint32_t q = 0; q < RB / CPB; q++) cp_async<CPB>(dst + q * CPB, src + q * CPB);
        }
        staged = true;
    }
    __syncthreads();
    const uint32_t b0 = s_boff[0], b1 = s_boff[1];
    cp_async_wait_all(); // (the siblings are read by the lane that copied them)
    if (b0 == b1) return;

    // the whole warp evaluates the Nb windows of one fired group
    auto eval_group = [&](uint32_t e_slot, uint64_t e_key, uint64_t e_g, uint32_t e_pos, uint32_t e_obase) {
        const unsigned char *e_tree = ff.tree + static_cast<size_t>(e_slot) * tree_stride;
        const uint64_t wm = batch_watermark(batch_off, batches, nbatches, e_pos);
        for (uint32_t i = lane; i < ff.nb; i += 32)
            ffat_eval_window<P>(ff, e_tree, e_key, e_g * ff.nb + i, wm, e_obase + i, out_res, out_ts, out_cap, prm);
    };

    // ---- consumer: one round. Every lane: fold the record it asked for ST_FLY rounds ago, pop the next position of its key, ask for it --
    uint32_t q_head = 0, q_tail = 0;     // my key's queue (indices grow; entry i lives at i % ST_QCAP)
    uint32_t total = 0;                  // items queued in the whole warp (warp-uniform)
    uint32_t fly_pos0 = 0, fly_pos1 = 0, fly_pos2 = 0; // positions of my gathers in flight, oldest first
    uint32_t fly_valid = 0;              // bit i: fly_pos<i> is a real gather
    uint32_t round_no = 0;               // warp-uniform: the stage slot of this round is round_no % ST_FLY
    static_assert(ST_FLY == 3, "three gathers in flight per lane");
    unsigned char *const my_stage = s_stage + static_cast<size_t>(tid) * ST_FLY * RB;
    uint32_t *const my_q = &q_pos[tid][0];
    auto consume_round = [&]() {
        cp_async_wait_group<ST_FLY - 1>(); // the copy committed ST_FLY rounds ago has landed (one group per round, empty or not)
        unsigned char *slot = my_stage + (round_no % ST_FLY) * RB;
        const bool folded = (fly_valid & 1u) != 0;
        const uint32_t it_pos = fly_pos0;
        if (folded) {
            alignas(16) R it;
            ld_rec<R>(slot, it);
            if (cp == 0) acc = it; else P::comb(acc, it, acc, prm);
            cp++; cons++;
        }
        fly_pos0 = fly_pos1; fly_pos1 = fly_pos2; fly_valid >>= 1;
        const bool popped = q_head != q_tail;
        if (popped) {
            const uint32_t pos = my_q[q_head % ST_QCAP];
            q_head++;
            const unsigned char *src = lifted + static_cast<size_t>(pos) * RB;
#pragma unroll
            for (uint32_t q = 0; q < RB / CPB; q++) cp_async<CPB>(slot + q * CPB, src + q * CPB);
            fly_pos2 = pos; fly_valid |= 1u << (ST_FLY - 1);
        }
        cp_async_commit();
        round_no++;
        total -= __popc(__ballot_sync(FULL, popped));
        if (!__any_sync(FULL, folded && cp == P32)) return;
        // ---- some key completed a pane (about once per round): leaf, root path, fired group --------------------------------------------------
        const bool completes = folded && cp == P32;
        // a key about to overwrite ring leaves while a group of it is still deferred: evaluate that group first
        uint32_t ev = __ballot_sync(FULL, completes && pend != ST_NONE);
        while (ev) {
            const int src = __ffs(ev) - 1;
            ev &= ev - 1;
            const uint32_t ti = __shfl_sync(FULL, pend, src);
            const Trigger tr = ff.trig[ti];
            eval_group(tr.slot, tr.key, tr.g, tr.last_pos, tr.obase);
            if (lane == static_cast<uint32_t>(src)) { ff.trig[ti].slot = INVALID_SLOT; pend = ST_NONE; } // k_ffat_windows skips voided entries
            __syncwarp();
        }
        bool eval_now = false;
        uint32_t ev_obase = 0; uint64_t ev_g = 0;
        if (completes) {
            cp = 0;
            const uint32_t lf = leaf;
            leaf = (leaf + 1) & (n - 1);
            st_rec<R>(my_tree + static_cast<size_t>(lf) * RB, acc);
            alignas(16) R cur = acc;
            for (uint32_t l = 0; l < (ff.lazy ? 0u : logn); l++) { // (lazy: only the leaf is written)
                alignas(16) R sb;
                if (staged && l < SIBL) ld_rec<R>(s_sib + (my_k * SIBL + l) * RB, sb);
                else ld_rec<R>(my_tree + static_cast<size_t>(level_off(n, l) + ((lf >> l) ^ 1u)) * RB, sb);
                alignas(16) R parent = cur;
                if ((lf >> l) & 1u) P::comb(sb, cur, parent, prm); else P::comb(cur, sb, parent, prm);
                cur = parent;
                st_rec<R>(my_tree + static_cast<size_t>(level_off(n, l + 1) + (lf >> (l + 1))) * RB, cur);
            }
            staged = false; // the next pane of the key reads the tree this one has just written
            if (cons == tt) { // the group fires: Nb windows, evaluated after the update (k_ffat_windows)
                const uint32_t obase = atomicAdd(n_out, ff.nb);
                const uint32_t ti = atomicAdd(ff.n_trig, 1u);
                if (ti < ff.trig_cap) {
                    Trigger tr; tr.key = key_of_slot(ff, my_slot); tr.g = g; tr.slot = my_slot; tr.last_pos = it_pos; tr.obase = obase; tr.pad = 0;
                    ff.trig[ti] = tr;
                    pend = ti;
                } else { eval_now = true; ev_obase = obase; ev_g = g; } // list full: evaluate here
                g++; tt += group_items;
            }
        }
        uint32_t pd = __ballot_sync(FULL, eval_now);
        while (pd) {
            const int src = __ffs(pd) - 1;
            pd &= pd - 1;
            const uint32_t e_obase = __shfl_sync(FULL, ev_obase, src), e_pos = __shfl_sync(FULL, it_pos, src);
            const uint64_t e_g = __shfl_sync(FULL, ev_g, src);
            const uint32_t e_slot = key_lo + warp * 32u + static_cast<uint32_t>(src);
            __syncwarp(); // the path nodes the source lane has just written
            eval_group(e_slot, key_of_slot(ff, e_slot), e_g, e_pos, e_obase);
            __syncwarp();
        }
    };

    // ---- the stream as ONE loop with one producer site and one consumer site (the consumer body is long: instantiating it at several call
    // sites costs more in instruction-cache misses than it saves in branches) ---------------------------------------------------------
    // producer: the bucket's pairs, 32 per step (loaded 128 at a time, one step ahead), into the per-key queues
    uint32_t sl0, sl1, sl2, sl3, ps0, ps1, ps2, ps3;     // the 128 pairs being queued
    uint32_t nsl0, nsl1, nsl2, nsl3, nps0, nps1, nps2, nps3; // the next 128, in flight
    auto load4 = [&](uint32_t base, uint32_t &a0, uint32_t &a1, uint32_t &a2, uint32_t &a3, uint32_t &p0, uint32_t &p1, uint32_t &p2, uint32_t &p3) {
        const uint32_t i0 = base + lane, i1 = i0 + 32, i2 = i0 + 64, i3 = i0 + 96;
        a0 = i0 < b1 ? bk_slots[i0] : INVALID_SLOT; a1 = i1 < b1 ? bk_slots[i1] : INVALID_SLOT;
        a2 = i2 < b1 ? bk_slots[i2] : INVALID_SLOT; a3 = i3 < b1 ? bk_slots[i3] : INVALID_SLOT;
        p0 = i0 < b1 ? bk_pos[i0] : 0u; p1 = i1 < b1 ? bk_pos[i1] : 0u; p2 = i2 < b1 ? bk_pos[i2] : 0u; p3 = i3 < b1 ? bk_pos[i3] : 0u;
    };
    load4(b0, nsl0, nsl1, nsl2, nsl3, nps0, nps1, nps2, nps3);
    sl0 = sl1 = sl2 = sl3 = INVALID_SLOT; ps0 = ps1 = ps2 = ps3 = 0;
    uint32_t next_base = b0;  // first pair of the 128 in nsl/nps
    uint32_t u = 4;           // group of the current 128 to queue next (4: fetch the next 128)
    bool stream_done = false; // every pair of the bucket has been queued
#pragma unroll 1
    for (;;) {
        bool want_round = total >= ST_BACKLOG || (stream_done && (total != 0 || __any_sync(FULL, fly_valid != 0)));
        if (!want_round) {
            if (stream_done) break;
            if (u == 4) { // the 128 pairs loaded a step ago become current; the following 128 start to fly
                if (next_base >= b1) { stream_done = true; continue; }
                sl0 = nsl0; sl1 = nsl1; sl2 = nsl2; sl3 = nsl3; ps0 = nps0; ps1 = nps1; ps2 = nps2; ps3 = nps3;
                next_base += 128;
                load4(next_base, nsl0, nsl1, nsl2, nsl3, nps0, nps1, nps2, nps3);
                u = 0;
            }
            const uint32_t slu = u == 0 ? sl0 : (u == 1 ? sl1 : (u == 2 ? sl2 : sl3));
            const uint32_t psu = u == 0 ? ps0 : (u == 1 ? ps1 : (u == 2 ? ps2 : ps3));
            const uint32_t lkf = slu - key_lo; // (slots outside the bucket's keys, the padding included, wrap to large values)
            const bool mine = lkf < kpc && (lkf >> 5) == warp;
            const uint32_t lk = lkf & 31u;
            const uint32_t mmask = __ballot_sync(FULL, mine);
            if (mmask == 0) { u++; continue; }
            // key-lane view: the group's items of MY key (five ballots transpose the item keys into one mask per key lane)
            uint32_t m = mmask;
#pragma unroll
            for (uint32_t b = 0; b < 5; b++) { const uint32_t B = __ballot_sync(FULL, mine && ((lk >> b) & 1u)); m &= ((lane >> b) & 1u) ? B : ~B; }
            const uint32_t incoming = __popc(m);
            want_round = __any_sync(FULL, q_tail - q_head + incoming > ST_QCAP); // (a key that floods its queue: drain a round, then retry this group)
            if (!want_round) {
                // item-lane view: my rank among the group's items of my key, my key's tail
                const uint32_t peers = __match_any_sync(FULL, mine ? lk : 32u + lane);
                const uint32_t tail = __shfl_sync(FULL, q_tail, lk);
                if (mine) q_pos[warp * 32u + lk][(tail + __popc(peers & lanemask_lt())) % ST_QCAP] = psu;
                q_tail += incoming;
                total += __popc(mmask);
                u++;
                __syncwarp(); // the queue entries are visible to the key lanes
                continue;
            }
        }
        consume_round();
    }
    cp_async_wait_all();
    // ---- my key's state back -----------------------------------------------------------------------------------------------------------
    if (has_key && cons != 0) {
        ff.cnt[my_slot] = st_c + cons;
        if (cp) st_rec<R>(ff.acc + static_cast<size_t>(my_slot) * RB, acc);
    }
}

// ------------------------------------------------------------------------------------------------------
// k_ffat_update: one warp per key that received items in this stream segment.
//   items of the key, in arrival order: lifted[sorted_pos[seg_off[slot] .. +seg_cnt[slot])] (gather = 1), or, when the last
//   sort pass also moved the records (gather = 0), lifted[seg_off[slot] .. +seg_cnt[slot])
//   -> ordered warp fold into the open pane (pane = gcd(win, slide) items)
//   -> completed pane = new FlatFAT leaf (ring of n_leaves panes) + recompute of its root path
//   -> when the key's count reaches the trigger: Nb window queries (greedy aligned-node fold, the same walk as
//      Compute_Results_Kernel, wf/flatfat_gpu.hpp:93-139, over panes instead of tuples)
// Window / trigger bookkeeping restates Ffat_Replica_GPU::process_wins_cb (wf/ffat_replica_gpu.hpp:830-867):
// groups fired so far G(c) = c < B ? 0 : 1 + (c - B) / (S*Nb); next_gwid = G*Nb; trigger = B + G*S*Nb.
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t level_off(uint32_t n_leaves, uint32_t level) { return 2u * n_leaves - ((2u * n_leaves) >> level); }

// watermark of the batch that holds compact position `pos` (batch_off has nbatches+1 ascending entries)
__device__ __forceinline__ uint64_t batch_watermark(const uint32_t *__restrict__ batch_off, const DevBatch *__restrict__ batches,
                                                    uint32_t nbatches, uint32_t pos)
{
    uint32_t lo = 0, hi = nbatches - 1;
    while (lo < hi) { const uint32_t mid = (lo + hi + 1) >> 1; if (batch_off[mid] <= pos) lo = mid; else hi = mid - 1; }
    return batches[lo].watermark;
}

// one window: result_t(key, gwid) folded left to right over the largest aligned FlatFAT nodes covering panes
// [gwid*sp, gwid*sp + wp) of the key's ring (the walk of Compute_Results_Kernel, wf/flatfat_gpu.hpp:109-136)
template <class P>
__device__ __forceinline__ void ffat_eval_window(const FfatDev &ff, const unsigned char *tree, uint64_t key, uint64_t gwid, uint64_t wm,
                                                 uint32_t opos, unsigned char *__restrict__ out_res, uint64_t *__restrict__ out_ts,
                                                 uint32_t out_cap, const typename P::params_t &prm)
{
    using R = typename P::result_t;
    constexpr uint32_t RB = sizeof(R);
    const uint32_t n = ff.n_leaves;
    alignas(16) R res = P::make_result(key, gwid, prm);
    uint32_t ws = static_cast<uint32_t>((gwid * ff.sp) & (n - 1));
    uint32_t remaining = ff.wp;
    if (ff.lazy) { // only the leaves are kept in global memory: fold them in order (the rare in-kernel evaluations; the deferred groups go
                   // through k_ffat_windows_lazy, which builds the levels on chip)
        for (; remaining > 0; remaining--) {
            alignas(16) R node;
            ld_rec<R>(tree + static_cast<size_t>(ws) * RB, node);
            P::comb(res, node, res, prm);
            ws = (ws + 1) & (n - 1);
        }
    }
    while (remaining > 0) {
        uint32_t range = (ws == 0) ? n : (ws & (0u - ws));
        const uint32_t pw = 1u << (31 - __clz(remaining));
        range = min(range, pw);
        const uint32_t level = 31 - __clz(range);
        alignas(16) R node;
        ld_rec<R>(tree + static_cast<size_t>(level_off(n, level) + (ws >> level)) * RB, node);
        P::comb(res, node, res, prm);
        ws = (ws + range) & (n - 1);
        remaining -= range;
    }
    if (opos < out_cap) {
        st_rec<R>(out_res + static_cast<size_t>(opos) * RB, res);
        if (out_ts != nullptr) out_ts[opos] = wm;
    } else atomicOr(ff.err_flags, 2u);
}

// deferred window groups: one thread per window
template <class P>
__global__ void __launch_bounds__(256) k_ffat_windows(const FfatDev ff, const uint32_t *__restrict__ batch_off,
                                                      const DevBatch *__restrict__ batches, uint32_t nbatches,
                                                      unsigned char *__restrict__ out_res, uint64_t *__restrict__ out_ts, uint32_t out_cap,
                                                      const typename P::params_t prm, uint32_t *__restrict__ n_out)
{
    using R = typename P::result_t;
    // the update kernels reserve output slots with atomicAdd(n_out, Nb) whether they fit or not: a count beyond the capacity is clamped
    // here (the last kernel of the call) and flagged, so that *n_out is always the number of results actually written
    if (blockIdx.x == 0 && threadIdx.x == 0 && n_out != nullptr) {
        if (*n_out > out_cap) { *n_out = out_cap; atomicOr(ff.err_flags, 2u); }
        if (ff.results_total != nullptr) *ff.results_total += *n_out;
    }
    const uint32_t nt = min(*ff.n_trig, ff.trig_cap);
    const uint64_t total = static_cast<uint64_t>(nt) * ff.nb;
    const size_t tree_stride = static_cast<size_t>(2 * ff.n_leaves - 1) * sizeof(R);
    for (uint64_t w = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; w < total; w += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        const uint32_t ti = static_cast<uint32_t>(w / ff.nb), i = static_cast<uint32_t>(w % ff.nb);
        const Trigger tr = ff.trig[ti];
        if (tr.slot == INVALID_SLOT) continue; // evaluated inside the update kernel (k_ffat_update_stream)
        const uint64_t wm = batch_watermark(batch_off, batches, nbatches, tr.last_pos);
        ffat_eval_window<P>(ff, ff.tree + static_cast<size_t>(tr.slot) * tree_stride, tr.key, tr.g * ff.nb + i, wm, tr.obase + i,
                            out_res, out_ts, out_cap, prm);
    }
}

// deferred window groups of a handle that keeps only the pane leaves (FfatDev::lazy): ONE WARP per fired group. The warp copies the key's
// n leaves to shared memory (coalesced), builds the n - 1 internal nodes level by level -- the nodes of a level are independent: lanes take
// them round-robin, one __syncwarp per level ("the FlatFAT levels with warp-level primitives") -- and then every lane evaluates its share of
// the group's Nb windows with the same greedy aligned-node walk over the on-chip tree. Dynamic shared memory: warps per block x 2 n x sizeof(R).
template <class P>
__global__ void __launch_bounds__(128) k_ffat_windows_lazy(const FfatDev ff, const uint32_t *__restrict__ batch_off,
                                                           const DevBatch *__restrict__ batches, uint32_t nbatches,
                                                           unsigned char *__restrict__ out_res, uint64_t *__restrict__ out_ts, uint32_t out_cap,
                                                           const typename P::params_t prm, uint32_t *__restrict__ n_out)
{
    using R = typename P::result_t;
    constexpr uint32_t RB = sizeof(R);
    extern __shared__ __align__(16) unsigned char lazy_smem[];
    if (blockIdx.x == 0 && threadIdx.x == 0 && n_out != nullptr) {
        if (*n_out > out_cap) { *n_out = out_cap; atomicOr(ff.err_flags, 2u); }
        if (ff.results_total != nullptr) *ff.results_total += *n_out;
    }
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const uint32_t n = ff.n_leaves, logn = ff.log_leaves;
    unsigned char *t = lazy_smem + static_cast<size_t>(warp) * 2u * n * RB; // (2 n - 1) nodes, levels back to back as in the global layout
    const uint32_t nt = min(*ff.n_trig, ff.trig_cap);
    const size_t tree_stride = static_cast<size_t>(2 * n - 1) * RB;
    FfatDev fs = ff; fs.lazy = 0; // the on-chip tree has every level: the ordinary walk
    for (uint32_t ti = blockIdx.x * wpb + warp; ti < nt; ti += gridDim.x * wpb) {
        const Trigger tr = ff.trig[ti];
        if (tr.slot == INVALID_SLOT) continue; // evaluated inside the update kernel
        const unsigned char *leaves = ff.tree + static_cast<size_t>(tr.slot) * tree_stride;
        for (uint32_t i = lane; i < n * (RB / 8); i += 32) reinterpret_cast<uint64_t *>(t)[i] = reinterpret_cast<const uint64_t *>(leaves)[i];
        __syncwarp();
        for (uint32_t l = 0; l < logn; l++) {
            const unsigned char *src = t + static_cast<size_t>(level_off(n, l)) * RB;
            unsigned char *dst = t + static_cast<size_t>(level_off(n, l + 1)) * RB;
            for (uint32_t i = lane; i < (n >> (l + 1)); i += 32) {
                alignas(16) R a, b, o;
                ld_rec<R>(src + static_cast<size_t>(2 * i) * RB, a); ld_rec<R>(src + static_cast<size_t>(2 * i + 1) * RB, b);
                o = a;
                P::comb(a, b, o, prm);
                st_rec<R>(dst + static_cast<size_t>(i) * RB, o);
            }
            __syncwarp();
        }
        const uint64_t wm = batch_watermark(batch_off, batches, nbatches, tr.last_pos);
        for (uint32_t i = lane; i < ff.nb; i += 32)
            ffat_eval_window<P>(fs, t, tr.key, tr.g * ff.nb + i, wm, tr.obase + i, out_res, out_ts, out_cap, prm);
        __syncwarp(); // the next group overwrites the on-chip tree
    }
}

template <class P>
__global__ void __launch_bounds__(256) k_ffat_update(const FfatDev ff, const unsigned char *__restrict__ lifted,
                                                     const uint32_t *__restrict__ sorted_pos,
                                                     const uint32_t *__restrict__ batch_off, const DevBatch *__restrict__ batches,
                                                     uint32_t nbatches, unsigned char *__restrict__ out_res,
                                                     uint64_t *__restrict__ out_ts, uint32_t out_cap, uint32_t *__restrict__ n_out,
                                                     uint32_t gather, const typename P::params_t prm, uint32_t use_heavy_list)
{
    using R = typename P::result_t;
    constexpr uint32_t RB = sizeof(R);
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t nslots = ff.dense ? ff.max_keys : min(*ff.n_slots, ff.max_keys);
    const uint32_t n = ff.n_leaves, logn = ff.log_leaves;
    const uint64_t P_ = ff.pane;
    const uint64_t group_items = ff.slide * ff.nb;
    const size_t tree_stride = static_cast<size_t>(2 * n - 1) * RB;

    // use_heavy_list = 1: only the keys k_ffat_update_lanes put on the heavy list; 0: every key of the segment
    const uint32_t nwork = use_heavy_list ? min(*ff.n_heavy, ff.max_keys) : nslots;
    for (uint32_t wi = gwarp; wi < nwork; wi += nwarps) {
        const uint32_t slot = use_heavy_list ? ff.heavy[wi] : wi;
        const uint32_t m = ff.seg_cnt[slot];
        if (m == 0) continue;
        const uint32_t off = ff.seg_off[slot];
        uint64_t c = ff.cnt[slot];
        const uint64_t key = key_of_slot(ff, slot);
        unsigned char *tree = ff.tree + static_cast<size_t>(slot) * tree_stride;
        alignas(16) R acc;
        if (c % P_ != 0) ld_rec<R>(ff.acc + static_cast<size_t>(slot) * RB, acc); // every lane keeps a copy
        uint64_t g = (c < ff.B) ? 0 : 1 + (c - ff.B) / group_items;
        uint64_t trig = ff.B + g * group_items;
        if (!gather) { // pull the key's (contiguous) records towards L2 before the chunk loop needs them
            const unsigned char *seg = lifted + static_cast<size_t>(off) * RB;
            const uint32_t lines = (m * RB + 127u) / 128u;
            for (uint32_t l = lane; l < lines && l < 128u; l += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(seg + static_cast<size_t>(l) * 128u));
        }

        uint32_t j = 0;
        while (j < m) {
            const uint32_t room = static_cast<uint32_t>(P_ - (c % P_));
            const uint32_t take = min(min(32u, m - j), room);
            alignas(16) R r;
            if (lane < take) {
                const uint32_t p = gather ? sorted_pos[off + j + lane] : (off + j + lane);
                ld_rec<R>(lifted + static_cast<size_t>(p) * RB, r);
            }
            // ordered fold: after the step with stride o, lane l holds items [l, l+2o) (clipped to take)
#pragma unroll
            for (uint32_t o = 1; o < 32; o <<= 1) {
                const R other = shfl_down_rec<R>(r, o);
                if (lane + o < take) P::comb(r, other, r, prm);
            }
            r = shfl_rec<R>(r, 0);
            if (c % P_ == 0) acc = r; else P::comb(acc, r, acc, prm);
            c += take; j += take;

            if (c % P_ == 0) { // pane complete -> leaf + root path
                const uint32_t leaf = static_cast<uint32_t>((c / P_ - 1) & (n - 1));
                alignas(16) R sib;
                const uint32_t plev = ff.lazy ? 0u : logn; // (lazy: only the leaf is written)
                if (lane < plev) { // lane l fetches the sibling of the path node at level l
                    const uint32_t idx = (leaf >> lane) ^ 1u;
                    ld_rec<R>(tree + static_cast<size_t>(level_off(n, lane) + idx) * RB, sib);
                }
                alignas(16) R cur = acc;
                if (lane == 0) st_rec<R>(tree + static_cast<size_t>(leaf) * RB, cur);
                for (uint32_t l = 0; l < plev; l++) {
                    const R s = shfl_rec<R>(sib, l);
                    alignas(16) R parent = cur; // key/id fields are don't-care in internal nodes
                    if ((leaf >> l) & 1u) P::comb(s, cur, parent, prm); else P::comb(cur, s, parent, prm);
                    cur = parent;
                    if (lane == 0) st_rec<R>(tree + static_cast<size_t>(level_off(n, l + 1) + (leaf >> (l + 1))) * RB, cur);
                }
                __syncwarp();

                if (c == trig) { // fire Nb windows: gwid = g*Nb + i
                    const uint32_t last_pos = sorted_pos[off + j - 1]; // arrival position of the triggering item
                    uint32_t obase = 0;
                    if (lane == 0) obase = atomicAdd(n_out, ff.nb);
                    obase = __shfl_sync(FULL, obase, 0);
                    // No further pane of this key can complete in this segment => the tree stays as it is now and the
                    // queries can run later, thread-per-window, in k_ffat_windows; otherwise evaluate them here.
                    bool deferred = (m - j) < ff.defer_items;
                    if (deferred) {
                        uint32_t ti = 0;
                        if (lane == 0) ti = atomicAdd(ff.n_trig, 1u);
                        ti = __shfl_sync(FULL, ti, 0);
                        if (ti < ff.trig_cap) {
                            if (lane == 0) { Trigger tr; tr.key = key; tr.g = g; tr.slot = slot; tr.last_pos = last_pos; tr.obase = obase; tr.pad = 0; ff.trig[ti] = tr; }
                        } else deferred = false;
                    }
                    if (!deferred) {
                        const uint64_t wm = batch_watermark(batch_off, batches, nbatches, last_pos);
                        for (uint32_t i = lane; i < ff.nb; i += 32)
                            ffat_eval_window<P>(ff, tree, key, g * ff.nb + i, wm, obase + i, out_res, out_ts, out_cap, prm);
                    }
                    g++; trig += group_items;
                    __syncwarp();
                }
            }
        }
        if (lane == 0) {
            ff.cnt[slot] = c;
            if (c % P_ != 0) st_rec<R>(ff.acc + static_cast<size_t>(slot) * RB, acc);
            ff.seg_cnt[slot] = 0;
            ff.seg_off[slot] = 0xffffffffu;
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// Per-batch keyed operators built on the sort: KeyBy_Emitter_GPU grouping, Reduce_GPU, key -> shard partition
// ------------------------------------------------------------------------------------------------------
// keys[i] = key_extr(tuple_i)   (Extract_Keys_Kernel wf/reduce_gpu.hpp:75-86, Extract_Dests_Kernel wf/keyby_emitter_gpu.hpp:68-81)
template <class P>
__global__ void k_extract_keys(const unsigned char *__restrict__ tuples, uint32_t n, uint64_t *__restrict__ keys,
                               uint32_t *__restrict__ dest, uint32_t num_shards, const typename P::params_t prm)
{
    using T = typename P::tuple_t;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const T *t = reinterpret_cast<const T *>(tuples + static_cast<size_t>(i) * sizeof(T));
        const uint64_t k = P::key(*t, prm);
        if (keys) keys[i] = k;
        if (dest) dest[i] = static_cast<uint32_t>(k % num_shards); // wf/keyby_emitter.hpp:215-217
    }
}

// head[i] = 1 when sorted position i starts a new key; map_idxs links equal neighbours
// (Compute_Mapping_Kernel, wf/keyby_emitter_gpu.hpp:84-100)
static __global__ void k_seg_heads(const uint64_t *__restrict__ skeys, const uint32_t *__restrict__ sidx, uint32_t n,
                            uint32_t *__restrict__ head, int32_t *__restrict__ map_idxs)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        head[i] = (i == 0 || skeys[i] != skeys[i - 1]) ? 1u : 0u;
        if (map_idxs) map_idxs[sidx[i]] = (i + 1 < n && skeys[i] == skeys[i + 1]) ? static_cast<int32_t>(sidx[i + 1]) : -1;
    }
}

// after the exclusive scan of head[] (seg[i] = index of the segment sorted position i belongs to, for heads):
// start_idxs[k] / dist_keys[k] of the k-th distinct key (unique_by_key_copy, wf/keyby_emitter_gpu.hpp:559-564),
// seg_begin[k] = first sorted position of segment k (used by the reduce), *n_keys = number of segments
static __global__ void k_seg_finish(const uint64_t *__restrict__ skeys, const uint32_t *__restrict__ sidx, const uint32_t *__restrict__ head_scan,
                             uint32_t n, int32_t *__restrict__ start_idxs, uint64_t *__restrict__ dist_keys,
                             uint32_t *__restrict__ seg_begin, uint32_t *__restrict__ n_keys)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const bool is_head = (i == 0 || skeys[i] != skeys[i - 1]);
        if (is_head) {
            const uint32_t k = head_scan[i];
            if (start_idxs) start_idxs[k] = static_cast<int32_t>(sidx[i]);
            if (dist_keys) dist_keys[k] = skeys[i];
            if (seg_begin) seg_begin[k] = i;
        }
        if (i == n - 1) {
            const uint32_t total = head_scan[i] + (is_head ? 1u : 0u); // exclusive scan value + own flag
            if (n_keys) *n_keys = total;
            if (seg_begin) seg_begin[total] = n;
        }
    }
}

// Reduce_GPU keyed: one warp per distinct key folds the key's items in arrival order with P::reduce, ts = max
// (thrust_reduce_func_gpu_t, wf/reduce_gpu.hpp:88-105; reduce_by_key :245-252). Output k = k-th smallest key.
template <class P>
__global__ void __launch_bounds__(256) k_reduce_segments(const unsigned char *__restrict__ tuples, const uint64_t *__restrict__ ts,
                                                         const uint32_t *__restrict__ sidx, const uint32_t *__restrict__ seg_begin,
                                                         const uint32_t *__restrict__ n_keys, unsigned char *__restrict__ out_tuples,
                                                         uint64_t *__restrict__ out_ts, const typename P::params_t prm)
{
    using T = typename P::tuple_t;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t nk = *n_keys;
    for (uint32_t k = gwarp; k < nk; k += nwarps) {
        const uint32_t b = seg_begin[k], e = seg_begin[k + 1];
        alignas(16) T acc; uint64_t mts = 0; bool have = false;
        for (uint32_t j = b; j < e; j += 32) {
            const uint32_t take = min(32u, e - j);
            alignas(16) T t; uint64_t tt = 0;
            if (lane < take) {
                const uint32_t i = sidx[j + lane];
                ld_rec<T>(tuples + static_cast<size_t>(i) * sizeof(T), t);
                tt = ts ? ts[i] : 0;
            }
#pragma unroll
            for (uint32_t o = 1; o < 32; o <<= 1) {
                const T other = shfl_down_rec<T>(t, o);
                const uint64_t ots = __shfl_down_sync(FULL, tt, o);
                if (lane + o < take) { t = P::reduce(t, other, prm); tt = tt < ots ? ots : tt; }
            }
            t = shfl_rec<T>(t, 0); tt = __shfl_sync(FULL, tt, 0);
            if (!have) { acc = t; mts = tt; have = true; }
            else { acc = P::reduce(acc, t, prm); mts = mts < tt ? tt : mts; }
        }
        if (lane == 0) {
            st_rec<T>(out_tuples + static_cast<size_t>(k) * sizeof(T), acc);
            if (out_ts) out_ts[k] = mts;
        }
    }
}

// ---- Reduce_GPU over K queued batches in one launch sequence: composite sort key (batch index << key_bits) | key ----------
// batch of global element index gi (boff has nb+1 ascending entries)
__device__ __forceinline__ uint32_t batch_of(const uint32_t *__restrict__ boff, uint32_t nb, uint32_t gi)
{
    uint32_t lo = 0, hi = nb - 1;
    while (lo < hi) { const uint32_t mid = (lo + hi + 1) >> 1; if (boff[mid] <= gi) lo = mid; else hi = mid - 1; }
    return lo;
}

template <class P>
__global__ void k_extract_keys_batches(const DevBatch *__restrict__ batches, const uint32_t *__restrict__ boff, uint32_t nb, uint32_t total,
                                       uint32_t key_bits, uint64_t *__restrict__ keys, const typename P::params_t prm)
{
    using T = typename P::tuple_t;
    const uint64_t mask = key_bits >= 64 ? ~0ull : ((1ull << key_bits) - 1ull);
    for (uint32_t gi = blockIdx.x * blockDim.x + threadIdx.x; gi < total; gi += gridDim.x * blockDim.x) {
        const uint32_t b = batch_of(boff, nb, gi);
        const T *t = reinterpret_cast<const T *>(batches[b].tuples + static_cast<size_t>(gi - boff[b]) * sizeof(T));
        keys[gi] = (key_bits >= 64 ? 0ull : (static_cast<uint64_t>(b) << key_bits)) | (P::key(*t, prm) & mask);
    }
}

constexpr uint32_t SEGT = 2048; // sorted positions per tile of the head count / finish kernels (256 threads x 8)

// counts[tile] = segment heads (sorted position whose key differs from its predecessor) in the tile
static __global__ void __launch_bounds__(256) k_head_tile_counts(const uint64_t *__restrict__ skeys, uint32_t n, uint32_t *__restrict__ counts)
{
    __shared__ uint32_t wsum[8];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, base = blockIdx.x * SEGT;
    uint32_t c = 0;
#pragma unroll
    for (uint32_t r = 0; r < SEGT / 256; r++) {
        const uint32_t i = base + r * 256 + tid;
        if (i < n && (i == 0 || skeys[i] != skeys[i - 1])) c++;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(FULL, c, o);
    if (lane == 0) wsum[warp] = c;
    __syncthreads();
    if (tid == 0) { uint32_t t = 0; for (int w = 0; w < 8; w++) t += wsum[w]; counts[blockIdx.x] = t; }
}

// tile_base = exclusive scan of the tile counts. seg_begin[k] = first sorted position of the k-th segment (global numbering),
// first_seg[b] = number of the first segment of batch b (0xffffffff when the batch has none), *n_segs = total, seg_begin[total] = n
static __global__ void __launch_bounds__(256) k_seg_finish_batches(const uint64_t *__restrict__ skeys, uint32_t n, uint32_t key_bits,
                                                            const uint32_t *__restrict__ tile_base, uint32_t *__restrict__ seg_begin,
                                                            uint32_t *__restrict__ first_seg, uint32_t *__restrict__ n_segs)
{
    __shared__ uint32_t wsum[8];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t first = blockIdx.x * SEGT + tid * (SEGT / 256); // thread t owns 8 consecutive positions
    bool head[SEGT / 256];
    uint32_t mine = 0;
#pragma unroll
    for (uint32_t r = 0; r < SEGT / 256; r++) {
        const uint32_t i = first + r;
        head[r] = i < n && (i == 0 || skeys[i] != skeys[i - 1]);
        mine += head[r] ? 1u : 0u;
    }
    uint32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(FULL, incl, o); if (lane >= static_cast<uint32_t>(o)) incl += v; }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    uint32_t k = tile_base[blockIdx.x] + incl - mine;
    for (uint32_t w = 0; w < warp; w++) k += wsum[w];
#pragma unroll
    for (uint32_t r = 0; r < SEGT / 256; r++) {
        const uint32_t i = first + r;
        if (head[r]) {
            seg_begin[k] = i;
            const uint64_t b = key_bits >= 64 ? 0ull : (skeys[i] >> key_bits);
            if (i == 0 || (key_bits < 64 && (skeys[i - 1] >> key_bits) != b)) first_seg[b] = k;
            k++;
        }
        if (i == n - 1) { *n_segs = k; seg_begin[k] = n; }
    }
}

// n_out[b] = segments of batch b (first_seg[] is ascending over the batches that have segments)
static __global__ void k_batch_seg_counts(const uint32_t *__restrict__ first_seg, uint32_t nb, const uint32_t *__restrict__ n_segs,
                                          const DevBatch *__restrict__ batches)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        uint32_t next = *n_segs;
        for (uint32_t b = nb; b-- > 0;) {
            uint32_t c = 0;
            if (first_seg[b] != 0xffffffffu) { c = next - first_seg[b]; next = first_seg[b]; }
            if (batches[b].n_out != nullptr) *batches[b].n_out = c;
        }
    }
}

constexpr uint32_t RB_LONG = 48; // segments longer than this are folded by a warp (k_reduce_segments_batches), the others by one thread

// one THREAD per segment (most keys of a batch occur once or twice): sequential fold, 4 tuples in flight; long segments
// are put on a list for the warp kernel
template <class P>
__global__ void __launch_bounds__(128) k_reduce_segments_batches_short(const DevBatch *__restrict__ batches, const uint32_t *__restrict__ boff,
                                                                       const uint64_t *__restrict__ skeys, const uint32_t *__restrict__ sidx,
                                                                       const uint32_t *__restrict__ seg_begin, const uint32_t *__restrict__ first_seg,
                                                                       const uint32_t *__restrict__ n_segs, uint32_t key_bits,
                                                                       uint32_t *__restrict__ long_list, uint32_t *__restrict__ n_long,
                                                                       const typename P::params_t prm)
{
    using T = typename P::tuple_t;
    const uint32_t nk = *n_segs;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < nk; k += gridDim.x * blockDim.x) {
        const uint32_t sb = seg_begin[k], se = seg_begin[k + 1];
        if (se - sb > RB_LONG) { long_list[atomicAdd(n_long, 1u)] = k; continue; }
        const uint32_t b = key_bits >= 64 ? 0u : static_cast<uint32_t>(skeys[sb] >> key_bits);
        const unsigned char *tuples = batches[b].tuples;
        const uint64_t *ts = batches[b].ts;
        const uint32_t base = boff[b];
        alignas(16) T acc; uint64_t mts = 0;
        {
            const uint32_t i = sidx[sb] - base;
            ld_rec<T>(tuples + static_cast<size_t>(i) * sizeof(T), acc);
            mts = ts ? ts[i] : 0;
        }
        for (uint32_t j = sb + 1; j < se; j += 2) { // two tuples in flight
            const uint32_t i0 = sidx[j] - base, i1 = (j + 1 < se) ? sidx[j + 1] - base : i0;
            alignas(16) T t0, t1;
            ld_rec<T>(tuples + static_cast<size_t>(i0) * sizeof(T), t0);
            if (j + 1 < se) ld_rec<T>(tuples + static_cast<size_t>(i1) * sizeof(T), t1);
            const uint64_t s0 = ts ? ts[i0] : 0, s1 = (ts && j + 1 < se) ? ts[i1] : 0;
            acc = P::reduce(acc, t0, prm); mts = mts < s0 ? s0 : mts;
            if (j + 1 < se) { acc = P::reduce(acc, t1, prm); mts = mts < s1 ? s1 : mts; }
        }
        const uint32_t o = k - first_seg[b];
        st_rec<T>(batches[b].out + static_cast<size_t>(o) * sizeof(T), acc);
        if (batches[b].ts_out) batches[b].ts_out[o] = mts;
    }
}

// one warp per long segment (list filled by the kernel above; long_list == nullptr: every segment): ordered fold with
// P::reduce, ts = max; output = the batch's out buffers
template <class P>
__global__ void __launch_bounds__(256) k_reduce_segments_batches(const DevBatch *__restrict__ batches, const uint32_t *__restrict__ boff,
                                                                 const uint64_t *__restrict__ skeys, const uint32_t *__restrict__ sidx,
                                                                 const uint32_t *__restrict__ seg_begin, const uint32_t *__restrict__ first_seg,
                                                                 const uint32_t *__restrict__ n_segs, uint32_t key_bits,
                                                                 const uint32_t *__restrict__ long_list, const uint32_t *__restrict__ n_long,
                                                                 const typename P::params_t prm)
{
    using T = typename P::tuple_t;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t nk = long_list ? *n_long : *n_segs;
    for (uint32_t kk = gwarp; kk < nk; kk += nwarps) {
        const uint32_t k = long_list ? long_list[kk] : kk;
        const uint32_t sb = seg_begin[k], se = seg_begin[k + 1];
        const uint32_t b = key_bits >= 64 ? 0u : static_cast<uint32_t>(skeys[sb] >> key_bits);
        const DevBatch bt = batches[b];
        const uint32_t base = boff[b];
        alignas(16) T acc; uint64_t mts = 0; bool have = false;
        for (uint32_t j = sb; j < se; j += 32) {
            const uint32_t take = min(32u, se - j);
            alignas(16) T t; uint64_t tt = 0;
            if (lane < take) {
                const uint32_t i = sidx[j + lane] - base;
                ld_rec<T>(bt.tuples + static_cast<size_t>(i) * sizeof(T), t);
                tt = bt.ts ? bt.ts[i] : 0;
            }
#pragma unroll
            for (uint32_t o = 1; o < 32; o <<= 1) {
                const T other = shfl_down_rec<T>(t, o);
                const uint64_t ots = __shfl_down_sync(FULL, tt, o);
                if (lane + o < take) { t = P::reduce(t, other, prm); tt = tt < ots ? ots : tt; }
            }
            t = shfl_rec<T>(t, 0); tt = __shfl_sync(FULL, tt, 0);
            if (!have) { acc = t; mts = tt; have = true; }
            else { acc = P::reduce(acc, t, prm); mts = mts < tt ? tt : mts; }
        }
        if (lane == 0) {
            const uint32_t o = k - first_seg[b];
            st_rec<T>(bt.out + static_cast<size_t>(o) * sizeof(T), acc);
            if (bt.ts_out) bt.ts_out[o] = mts;
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// Time-based windows, front end (Ffat_Replica_GPU::process_batch_tb, wf/ffat_replica_gpu.hpp:870-1019; PendingPanes_Queue
// :263-420; Aggregate_Panes_Kernel :214-260). Per input batch: lift + (slot, pane) composite key -> stable sort ->
// per-(key, pane) partial in arrival order -> merged into the key's ring of pending panes -> for every key present in the
// batch, the panes of the groups the watermark has completed are popped, in pane order, into an array of lifted records
// that the count-based back end (a second handle over the "lifted" program, window / slide in panes) consumes as one
// batch: it fires exactly one group of Nb windows per popped group, with ts = the batch watermark.
// ------------------------------------------------------------------------------------------------------

struct TbDev {
    uint64_t pane_len, Bp, group;      // pane length (timestamp units), panes of the first group, panes of every further group
    uint32_t capq;                     // ring capacity per key (panes)
    uint32_t kbits;                    // this batch: sort key = (slot << kbits) | (pane - first pending pane of the key); 2^kbits - 1 = late
                                       // pane (already consumed), slot >= max_keys = no tuple (filtered out)
    uint64_t *first;                   // id of the first pending pane of every key
    uint32_t *num, *num_new;           // pending panes (before / after this batch)
    uint64_t *trig;                    // pane_id_triggerer
    uint32_t *done;                    // firstWinDone
    unsigned char *ring;               // max_keys x capq results, pane p of key s at (s * capq + p % capq)
    uint32_t *present, *n_present;     // slots of the keys of this batch
    uint32_t *cnt;                     // panes to pop per present key -> exclusive offsets
    uint32_t *ignored;                 // tuples older than the first incomplete pane (statistic)
    uint32_t *need;                    // ring capacity this batch needs: max over its tuples of pane - first pending pane + 2
    uint32_t *err;                     // bit 2: ring overflow / pane id out of range
};

template <class P>
__global__ void k_tb_lift(const unsigned char *__restrict__ tuples, const uint64_t *__restrict__ ts, uint32_t n, const FfatDev ff,
                          const TbDev tb, uint64_t first_incomplete, unsigned char *__restrict__ lifted, uint64_t *__restrict__ ckeys,
                          const typename P::params_t prm)
{
    using T = typename P::tuple_t;
    using R = typename P::result_t;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        alignas(16) T t;
        ld_rec<T>(tuples + static_cast<size_t>(i) * sizeof(T), t);
        P::map(t, prm);
        uint64_t ck = ~0ull;
        if (P::filter(t, prm)) {
            alignas(16) R r;
            P::lift(t, r, prm);
            st_rec<R>(lifted + static_cast<size_t>(i) * sizeof(R), r);
            const uint32_t slot = slot_of_key(ff, P::key(t, prm));
            const uint64_t pane = ts[i] / tb.pane_len;                 // Lifting_Kernel_TB_Keyed :164
            if (pane < first_incomplete) atomicAdd(tb.ignored, 1u);    // :165-167
            if (slot != INVALID_SLOT) { // raw key: slot, pane relative to the key's first pending pane (0xffffffff: older = late)
                const uint64_t f0 = tb.first[slot];
                uint32_t rel = 0xffffffffu;
                if (pane >= f0) {
                    if (pane - f0 >= 0xfffffff0ull) atomicOr(tb.err, 4u);
                    else { rel = static_cast<uint32_t>(pane - f0); atomicMax(tb.need, rel + 2u); } // push_panes :367-372
                }
                ck = (static_cast<uint64_t>(slot) << 32) | rel;
            }
        }
        ckeys[i] = ck;
    }
}

// sort keys of the batch once the number of pane bits it needs is known: (slot << kbits) | relative pane
static __global__ void k_tb_pack(uint64_t *__restrict__ ckeys, uint32_t n, uint32_t kbits, uint32_t invalid_slot)
{
    const uint64_t late = (1ull << kbits) - 1ull;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint64_t c = ckeys[i];
        if (c == ~0ull) { ckeys[i] = static_cast<uint64_t>(invalid_slot) << kbits; continue; }
        const uint32_t rel = static_cast<uint32_t>(c);
        ckeys[i] = ((c >> 32) << kbits) | (rel == 0xffffffffu ? late : static_cast<uint64_t>(rel));
    }
}

// partial of every (key, pane) of the batch: fold of the lifted results in arrival order (thrust::reduce_by_key :925-935)
template <class P>
__global__ void k_tb_reduce(const unsigned char *__restrict__ lifted, const uint64_t *__restrict__ skeys, const uint32_t *__restrict__ sidx,
                            const uint32_t *__restrict__ seg_begin, const uint32_t *__restrict__ n_segs, unsigned char *__restrict__ part,
                            uint32_t kbits, uint32_t max_keys, const typename P::params_t prm)
{
    using R = typename P::result_t;
    const uint32_t nk = *n_segs;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < nk; k += gridDim.x * blockDim.x) {
        const uint32_t sb = seg_begin[k], se = seg_begin[k + 1];
        if ((skeys[sb] >> kbits) >= max_keys || (skeys[sb] & ((1ull << kbits) - 1ull)) == (1ull << kbits) - 1ull) continue; // no tuple / late pane
        alignas(16) R acc;
        ld_rec<R>(lifted + static_cast<size_t>(sidx[sb]) * sizeof(R), acc);
        for (uint32_t j = sb + 1; j < se; j++) {
            alignas(16) R r;
            ld_rec<R>(lifted + static_cast<size_t>(sidx[j]) * sizeof(R), r);
            P::comb(acc, r, acc, prm);
        }
        st_rec<R>(part + static_cast<size_t>(k) * sizeof(R), acc);
    }
}

// partials (ascending slot, ascending pane) -> the keys' rings of pending panes; the last partial of a key records the
// new number of pending panes and lists the key as present
template <class P>
__global__ void k_tb_merge(const uint64_t *__restrict__ skeys, const uint32_t *__restrict__ seg_begin, const uint32_t *__restrict__ n_segs,
                           const unsigned char *__restrict__ part, const FfatDev ff, const TbDev tb, const typename P::params_t prm)
{
    using R = typename P::result_t;
    const uint32_t nk = *n_segs;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < nk; k += gridDim.x * blockDim.x) {
        const uint64_t ck = skeys[seg_begin[k]];
        const uint64_t relmask = (1ull << tb.kbits) - 1ull;
        if ((ck >> tb.kbits) >= ff.max_keys) continue;                // filtered tuples
        const uint32_t slot = static_cast<uint32_t>(ck >> tb.kbits);
        const uint64_t first_id = tb.first[slot];
        const bool late = (ck & relmask) == relmask;                  // pane already consumed (:232-234)
        const uint64_t pane = late ? 0 : first_id + (ck & relmask);
        const uint32_t num = tb.num[slot];
        const uint64_t have_end = first_id + num;                      // first pane id that is not in the ring yet
        unsigned char *ring = tb.ring + static_cast<size_t>(slot) * tb.capq * sizeof(R);
        const uint64_t ckn = (k + 1 < nk) ? skeys[seg_begin[k + 1]] : ~0ull;
        const bool last_of_key = (ckn >> tb.kbits) != slot; // (ckn = ~0 past the end: never a slot)
        bool stored = false;
        if (!late) {
            if (pane - first_id >= tb.capq) atomicOr(tb.err, 4u);
            else {
                alignas(16) R v;
                ld_rec<R>(part + static_cast<size_t>(k) * sizeof(R), v);
                unsigned char *dst = ring + (pane % tb.capq) * sizeof(R);
                if (pane < have_end) { alignas(16) R old; ld_rec<R>(dst, old); P::comb(old, v, old, prm); st_rec<R>(dst, old); } // :236
                else {
                    st_rec<R>(dst, v);
                    uint64_t lower = have_end;                          // missing panes below this one become empty panes (:239-257)
                    if (k > 0) {
                        const uint64_t ckp = skeys[seg_begin[k - 1]];
                        if ((ckp >> tb.kbits) == slot && first_id + (ckp & relmask) + 1 > lower) lower = first_id + (ckp & relmask) + 1;
                    }
                    const uint64_t key = key_of_slot(ff, slot);
                    for (uint64_t m = lower; m < pane; m++) {
                        alignas(16) R e = P::make_result(key, 0, prm);
                        st_rec<R>(ring + (m % tb.capq) * sizeof(R), e);
                    }
                }
                stored = true;
            }
        }
        if (last_of_key) {
            uint32_t nn = num;
            if (stored && pane >= have_end) nn = static_cast<uint32_t>(pane - first_id + 1);
            tb.num_new[slot] = nn;
            tb.present[atomicAdd(tb.n_present, 1u)] = slot;
        }
    }
}

// the rings grow (PendingPanes_Queue::resize :326-357): pending pane p of key s moves from p % old_cap to p % new_cap
static __global__ void k_tb_ring_resize(const TbDev tb, const unsigned char *__restrict__ old_ring, uint32_t old_cap, unsigned char *__restrict__ new_ring,
                                        uint32_t new_cap, uint32_t max_keys, uint32_t rbytes)
{
    for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < max_keys; s += gridDim.x * blockDim.x) {
        const uint64_t f0 = tb.first[s];
        const uint32_t num = tb.num[s];
        for (uint32_t j = 0; j < num; j++) {
            const uint64_t *src = reinterpret_cast<const uint64_t *>(old_ring + (static_cast<size_t>(s) * old_cap + (f0 + j) % old_cap) * rbytes);
            uint64_t *dst = reinterpret_cast<uint64_t *>(new_ring + (static_cast<size_t>(s) * new_cap + (f0 + j) % new_cap) * rbytes);
            for (uint32_t q = 0; q < rbytes / 8; q++) dst[q] = src[q];
        }
    }
}

// panes every present key pops now: groups completed by the watermark (process_wins_tb :1029-1046), first Bp then `group` each
static __global__ void k_tb_pop_count(const TbDev tb, uint64_t first_incomplete)
{
    const uint32_t np = *tb.n_present;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < np; i += gridDim.x * blockDim.x) {
        const uint32_t slot = tb.present[i];
        uint64_t trig = tb.trig[slot];
        bool done = tb.done[slot] != 0;
        uint64_t c = 0;
        while (trig < first_incomplete) { c += done ? tb.group : tb.Bp; done = true; trig += tb.group; }
        tb.cnt[i] = static_cast<uint32_t>(c > 0x7fffffffull ? 0x7fffffffull : c);
    }
}

// exclusive scan of the pop counts of the *n_present keys of the batch (one CTA), total to *total_out
static __global__ void __launch_bounds__(1024) k_tb_scan_present(uint32_t *__restrict__ cnt, const uint32_t *__restrict__ n_present, uint32_t *__restrict__ total_out)
{
    __shared__ uint32_t warp_sums[32];
    const uint32_t total = *n_present;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t per = (total + 1023) / 1024;
    const uint32_t begin = min(tid * per, total), end = min(begin + per, total);
    uint32_t sum = 0;
    for (uint32_t i = begin; i < end; i++) sum += cnt[i];
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(FULL, incl, o); if (lane >= static_cast<uint32_t>(o)) incl += v; }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = warp_sums[lane], wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(FULL, wi, o); if (lane >= static_cast<uint32_t>(o)) wi += v; }
        warp_sums[lane] = wi - w;
        if (lane == 31) *total_out = wi;
    }
    __syncthreads();
    uint32_t run = warp_sums[warp] + incl - sum;
    for (uint32_t i = begin; i < end; i++) { const uint32_t v = cnt[i]; cnt[i] = run; run += v; }
}

template <class P>
__global__ void k_tb_pop_write(const FfatDev ff, const TbDev tb, uint64_t first_incomplete, const uint32_t *__restrict__ offs,
                               unsigned char *__restrict__ popped, uint32_t *__restrict__ popped_slots, uint32_t popped_cap,
                               const typename P::params_t prm)
{
    using R = typename P::result_t;
    const uint32_t np = *tb.n_present;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < np; i += gridDim.x * blockDim.x) {
        const uint32_t slot = tb.present[i];
        uint64_t trig = tb.trig[slot], first_id = tb.first[slot];
        uint32_t num = tb.num_new[slot];
        bool done = tb.done[slot] != 0;
        const uint64_t key = key_of_slot(ff, slot);
        const unsigned char *ring = tb.ring + static_cast<size_t>(slot) * tb.capq * sizeof(R);
        uint32_t w = offs[i];
        while (trig < first_incomplete) {
            const uint64_t need = done ? tb.group : tb.Bp;
            for (uint64_t m = 0; m < need; m++, w++) {                  // pop_and_add :394-415; a missing pane is an empty pane
                alignas(16) R v;
                if (m < num) ld_rec<R>(ring + ((first_id + m) % tb.capq) * sizeof(R), v);
                else v = P::make_result(key, 0, prm);
                if (w < popped_cap) { st_rec<R>(popped + static_cast<size_t>(w) * sizeof(R), v); popped_slots[w] = slot; }
            }
            first_id += need; num = num > need ? static_cast<uint32_t>(num - need) : 0u;
            done = true; trig += tb.group;
        }
        tb.first[slot] = first_id; tb.num[slot] = num; tb.trig[slot] = trig; tb.done[slot] = done ? 1u : 0u;
    }
}

// ------------------------------------------------------------------------------------------------------
// Keyed-stateful Map_GPU / Filter_GPU (wf/map_gpu.hpp:80-102, :212-299; wf/filter_gpu.hpp:91-117, :247-355): the same
// shape as the window update -- slots of the segment's tuples, ONE wide partition pass into 1024 buckets of consecutive
// slots, one CTA per bucket splits its items by key (stable) and ONE THREAD per key walks its run in arrival order with
// the key's state in registers. Stateful filter: keep flags, then a stable per-batch compaction.
// ------------------------------------------------------------------------------------------------------
template <class P>
__global__ void k_ks_slots(const DevBatch *__restrict__ batches, const uint32_t *__restrict__ boff, uint32_t nb, uint32_t total, const FfatDev ff,
                           uint32_t *__restrict__ slots, const typename P::params_t prm)
{
    using T = typename P::tuple_t;
    for (uint32_t gi = blockIdx.x * blockDim.x + threadIdx.x; gi < total; gi += gridDim.x * blockDim.x) {
        const uint32_t b = batch_of(boff, nb, gi);
        const T *t = reinterpret_cast<const T *>(batches[b].tuples + static_cast<size_t>(gi - boff[b]) * sizeof(T));
        slots[gi] = slot_of_key(ff, P::key(*t, prm));
    }
}

constexpr uint32_t KS_KEYS = 64, KS_THREADS = 128, KS_IT = 18, KS_CAP = KS_THREADS * KS_IT;

template <class P, bool FILTER>
__global__ void __launch_bounds__(KS_THREADS) k_ks_apply(const FfatDev ff, const DevBatch *__restrict__ batches, const uint32_t *__restrict__ boff,
                                                         uint32_t nb, const uint32_t *__restrict__ bk_slots, const uint32_t *__restrict__ bk_pos,
                                                         const uint32_t *__restrict__ digit_counts, uint32_t shift, unsigned char *__restrict__ states,
                                                         unsigned char *__restrict__ keep, const typename P::params_t prm)
{
    using T = typename P::tuple_t;
    using S = typename P::state_t;
    constexpr uint32_t NW = KS_THREADS / 32, DPT = OSW_DIGITS / KS_THREADS, HS = KS_THREADS + 2;
    __shared__ uint32_t s_idx[KS_CAP];                 // positions (global tuple index in the segment), key-major
    __shared__ uint16_t hist[KS_KEYS][HS];             // private key counts of every thread -> exclusive over the threads
    __shared__ uint32_t htot[2][KS_KEYS], kcnt[KS_KEYS], koff[KS_KEYS];
    __shared__ uint32_t misc[NW], s_boff[2];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, bucket = blockIdx.x;
    const uint32_t kpc = min(KS_KEYS, 1u << shift), key_lo = bucket << shift;
    { // bucket range = exclusive scan of the pass histogram (as in k_ffat_update_buckets)
        uint32_t cc[DPT];
#pragma unroll
        for (uint32_t q = 0; q < DPT / 4; q++) {
            const uint4 v = reinterpret_cast<const uint4 *>(digit_counts)[tid * (DPT / 4) + q];
            cc[4 * q] = v.x; cc[4 * q + 1] = v.y; cc[4 * q + 2] = v.z; cc[4 * q + 3] = v.w;
        }
        uint32_t sum = 0;
#pragma unroll
        for (uint32_t q = 0; q < DPT; q++) sum += cc[q];
        uint32_t incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(FULL, incl, o); if (lane >= static_cast<uint32_t>(o)) incl += v; }
        if (lane == 31) misc[warp] = incl;
        __syncthreads();
        if (tid == bucket / DPT) {
            uint32_t base = incl - sum;
            for (uint32_t w = 0; w < warp; w++) base += misc[w];
            uint32_t own = 0;
#pragma unroll
            for (uint32_t q = 0; q < DPT; q++) { if (q < bucket % DPT) base += cc[q]; if (q == bucket % DPT) own = cc[q]; }
            s_boff[0] = base; s_boff[1] = base + own;
        }
        __syncthreads();
    }
    uint32_t cursor = s_boff[0];
    const uint32_t bend = s_boff[1];
    if (cursor == bend) return;
    // the key's state stays in registers across the chunks of the bucket
    alignas(8) S st;
    const bool has_key = tid < kpc && key_lo + tid < ff.max_keys;
    bool touched = false;
    if (has_key) st = *reinterpret_cast<const S *>(states + static_cast<size_t>(key_lo + tid) * sizeof(S));
    while (cursor < bend) {
        const uint32_t nsel = min(KS_CAP, bend - cursor);
        uint32_t ek[KS_IT], ep[KS_IT];
#pragma unroll
        for (uint32_t r = 0; r < KS_IT; r++) {
            const uint32_t i = tid * KS_IT + r;
            ek[r] = KS_KEYS; ep[r] = 0;
            if (i < nsel) {
                const uint32_t lk = bk_slots[cursor + i] - key_lo; // slots outside the bucket's keys (invalid slots) are dropped
                ep[r] = bk_pos[cursor + i];
                if (lk < kpc) ek[r] = lk;
            }
        }
        {
            uint32_t *z = reinterpret_cast<uint32_t *>(&hist[0][0]);
            for (uint32_t i = tid; i < KS_KEYS * HS / 2; i += KS_THREADS) z[i] = 0;
        }
        __syncthreads();
#pragma unroll
        for (uint32_t r = 0; r < KS_IT; r++) if (ek[r] < KS_KEYS) hist[ek[r]][tid]++;
        __syncthreads();
        {
            const uint32_t k = tid & 63u, half = tid >> 6;
            uint16_t *row = &hist[k][half * 64];
            uint32_t run = 0;
#pragma unroll 16
            for (uint32_t i = 0; i < 64; i++) { const uint32_t c = row[i]; row[i] = static_cast<uint16_t>(run); run += c; }
            htot[half][k] = run;
        }
        __syncthreads();
        if (warp == 0) {
            const uint32_t a0 = htot[0][lane] + htot[1][lane], a1 = htot[0][lane + 32] + htot[1][lane + 32];
            uint32_t i0 = a0, i1 = a1;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v0 = __shfl_up_sync(FULL, i0, o), v1 = __shfl_up_sync(FULL, i1, o);
                if (lane >= static_cast<uint32_t>(o)) { i0 += v0; i1 += v1; }
            }
            const uint32_t t0 = __shfl_sync(FULL, i0, 31);
            kcnt[lane] = a0; kcnt[lane + 32] = a1;
            koff[lane] = i0 - a0; koff[lane + 32] = t0 + i1 - a1;
        }
        __syncthreads();
        {
            const uint32_t hb = tid >> 6;
#pragma unroll
            for (uint32_t r = 0; r < KS_IT; r++) {
                const uint32_t k = ek[r];
                if (k < KS_KEYS) {
                    const uint32_t rank = hist[k][tid];
                    hist[k][tid] = static_cast<uint16_t>(rank + 1);
                    s_idx[koff[k] + (hb ? htot[0][k] : 0u) + rank] = ep[r];
                }
            }
        }
        __syncthreads();
        if (has_key) { // one thread per key: the run in arrival order
            const uint32_t m = kcnt[tid], off = koff[tid];
            for (uint32_t j = 0; j < m; j++) {
                const uint32_t gi = s_idx[off + j];
                const uint32_t b = batch_of(boff, nb, gi);
                unsigned char *tp = const_cast<unsigned char *>(batches[b].tuples) + static_cast<size_t>(gi - boff[b]) * sizeof(T);
                alignas(16) T t;
                ld_rec<T>(tp, t);
                if constexpr (FILTER) keep[gi] = P::filter_stateful(t, st, prm) ? 1 : 0;
                else P::map_stateful(t, st, prm);
                st_rec<T>(tp, t);
            }
            touched |= m != 0;
        }
        cursor += nsel;
        __syncthreads();
    }
    if (has_key && touched) *reinterpret_cast<S *>(states + static_cast<size_t>(key_lo + tid) * sizeof(S)) = st;
}

// stable per-batch compaction by the keep flags of a stateful filter: tile counts, scan, scatter
static __global__ void __launch_bounds__(256) k_flag_tile_counts(const unsigned char *__restrict__ keep, uint32_t n, uint32_t *__restrict__ counts)
{
    __shared__ uint32_t wsum[8];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, base = blockIdx.x * SEGT;
    uint32_t c = 0;
#pragma unroll
    for (uint32_t r = 0; r < SEGT / 256; r++) { const uint32_t i = base + r * 256 + tid; if (i < n && keep[i]) c++; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(FULL, c, o);
    if (lane == 0) wsum[warp] = c;
    __syncthreads();
    if (tid == 0) { uint32_t t = 0; for (int w = 0; w < 8; w++) t += wsum[w]; counts[blockIdx.x] = t; }
}

// rank_start[b] = survivors before batch b (global rank of its first tuple); n_out of every batch
static __global__ void k_flag_batch_starts(const unsigned char *__restrict__ keep, const uint32_t *__restrict__ tile_base, const uint32_t *__restrict__ boff,
                                           uint32_t nb, uint32_t n, uint32_t *__restrict__ rank_start, const DevBatch *__restrict__ batches)
{
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b > nb) return;
    const uint32_t gi = boff[b]; // boff[nb] = n
    uint32_t r;
    if (gi >= n) { // total survivors
        const uint32_t lt = (n - 1) / SEGT;
        r = tile_base[lt];
        for (uint32_t i = lt * SEGT; i < n; i++) r += keep[i] ? 1u : 0u;
    } else {
        const uint32_t t = gi / SEGT;
        r = tile_base[t];
        for (uint32_t i = t * SEGT; i < gi; i++) r += keep[i] ? 1u : 0u;
    }
    rank_start[b] = r;
}
static __global__ void k_flag_batch_counts(const uint32_t *__restrict__ rank_start, uint32_t nb, const DevBatch *__restrict__ batches)
{
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < nb && batches[b].n_out != nullptr) *batches[b].n_out = rank_start[b + 1] - rank_start[b];
}

template <class P>
__global__ void __launch_bounds__(256) k_flag_scatter(const unsigned char *__restrict__ keep, const uint32_t *__restrict__ tile_base,
                                                      const uint32_t *__restrict__ boff, uint32_t nb, uint32_t n,
                                                      const uint32_t *__restrict__ rank_start, const DevBatch *__restrict__ batches)
{
    using T = typename P::tuple_t;
    __shared__ uint32_t wsum[8];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t first = blockIdx.x * SEGT + tid * (SEGT / 256);
    bool k[SEGT / 256];
    uint32_t mine = 0;
#pragma unroll
    for (uint32_t r = 0; r < SEGT / 256; r++) { const uint32_t i = first + r; k[r] = i < n && keep[i]; mine += k[r] ? 1u : 0u; }
    uint32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(FULL, incl, o); if (lane >= static_cast<uint32_t>(o)) incl += v; }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    uint32_t rank = tile_base[blockIdx.x] + incl - mine;
    for (uint32_t w = 0; w < warp; w++) rank += wsum[w];
#pragma unroll
    for (uint32_t r = 0; r < SEGT / 256; r++) {
        if (k[r]) {
            const uint32_t gi = first + r, b = batch_of(boff, nb, gi);
            const DevBatch bt = batches[b];
            const uint32_t li = gi - boff[b], o = rank - rank_start[b];
            alignas(16) T t;
            ld_rec<T>(bt.tuples + static_cast<size_t>(li) * sizeof(T), t);
            st_rec<T>(bt.out + static_cast<size_t>(o) * sizeof(T), t);
            if (bt.ts_out) bt.ts_out[o] = bt.ts[li];
            rank++;
        }
    }
}

// Reduce_GPU un-keyed: the whole batch folded into one item, starting from a default-constructed item
// (thrust::reduce with init = batch_item_gpu_t<tuple_t>(), wf/reduce_gpu.hpp:264-273). One CTA of 1024 threads.
template <class P>
__global__ void __launch_bounds__(1024) k_reduce_all(const unsigned char *__restrict__ tuples, const uint64_t *__restrict__ ts, uint32_t n,
                                                     unsigned char *__restrict__ out_tuple, uint64_t *__restrict__ out_ts,
                                                     const typename P::params_t prm)
{
    using T = typename P::tuple_t;
    __shared__ __align__(16) unsigned char sm[32 * sizeof(T)];
    __shared__ uint64_t smts[32];
    __shared__ uint32_t smhave[32];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // thread t folds the contiguous chunk [t*per, (t+1)*per): order preserved
    const uint32_t per = (n + 1023) / 1024;
    const uint32_t b = min(tid * per, n), e = min(b + per, n);
    alignas(16) T acc; uint64_t mts = 0; bool have = false;
    for (uint32_t i = b; i < e; i++) {
        alignas(16) T t; ld_rec<T>(tuples + static_cast<size_t>(i) * sizeof(T), t);
        const uint64_t tt = ts ? ts[i] : 0;
        if (!have) { acc = t; mts = tt; have = true; } else { acc = P::reduce(acc, t, prm); mts = mts < tt ? tt : mts; }
    }
    // ordered combine across lanes, then across warps (a lane/warp without items is skipped)
#pragma unroll
    for (uint32_t o = 1; o < 32; o <<= 1) {
        const T other = shfl_down_rec<T>(acc, o);
        const uint64_t ots = __shfl_down_sync(FULL, mts, o);
        const bool ohave = __shfl_down_sync(FULL, have ? 1u : 0u, o) != 0;
        if (lane + o < 32 && ohave) {
            if (have) { acc = P::reduce(acc, other, prm); mts = mts < ots ? ots : mts; } else { acc = other; mts = ots; have = true; }
        }
    }
    if (lane == 0) { st_rec<T>(sm + warp * sizeof(T), acc); smts[warp] = mts; smhave[warp] = have ? 1u : 0u; }
    __syncthreads();
    if (tid == 0) {
        T init{};                 // default-constructed tuple, timestamp 0
        alignas(16) T r = init; uint64_t rts = 0;
        for (uint32_t w = 0; w < 32; w++) if (smhave[w]) {
            alignas(16) T t; ld_rec<T>(sm + w * sizeof(T), t);
            r = P::reduce(r, t, prm); rts = rts < smts[w] ? smts[w] : rts;
        }
        st_rec<T>(out_tuple, r);
        if (out_ts) *out_ts = rts;
    }
}

// payload gather after a stable partition: out[j] = in[perm[j]] (tuples and timestamps)
template <class P>
__global__ void k_gather_tuples(const unsigned char *__restrict__ tuples, const uint64_t *__restrict__ ts, const uint32_t *__restrict__ perm,
                                uint32_t n, unsigned char *__restrict__ out_tuples, uint64_t *__restrict__ out_ts)
{
    using T = typename P::tuple_t;
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        const uint32_t i = perm[j];
        alignas(16) T t; ld_rec<T>(tuples + static_cast<size_t>(i) * sizeof(T), t);
        st_rec<T>(out_tuples + static_cast<size_t>(j) * sizeof(T), t);
        if (ts && out_ts) out_ts[j] = ts[i];
    }
}

// seg_off[d] for d in [0, num_shards]: first sorted position whose destination is >= d (sorted dest array)
static __global__ void k_shard_offsets(const uint32_t *__restrict__ sdest, uint32_t n, uint32_t num_shards, uint32_t *__restrict__ seg_off)
{
    for (uint32_t d = blockIdx.x * blockDim.x + threadIdx.x; d <= num_shards; d += gridDim.x * blockDim.x) {
        uint32_t lo = 0, hi = n; // lower_bound(sdest, d)
        while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (sdest[mid] < d) lo = mid + 1; else hi = mid; }
        seg_off[d] = lo;
    }
}

// pipelined window operator: hand the results of a finished segment over to the caller's buffers
static __global__ void k_copy_results(const unsigned char *__restrict__ src, const uint64_t *__restrict__ src_ts, const uint32_t *__restrict__ src_n,
                               uint32_t rec_bytes, unsigned char *__restrict__ dst, uint64_t *__restrict__ dst_ts, uint32_t dst_cap,
                               uint32_t *__restrict__ n_out, uint32_t *__restrict__ err_flags)
{
    const uint32_t n = *src_n;
    const uint32_t m = min(n, dst_cap);
    const uint64_t words = static_cast<uint64_t>(m) * (rec_bytes / 8);
    const uint64_t *s8 = reinterpret_cast<const uint64_t *>(src);
    uint64_t *d8 = reinterpret_cast<uint64_t *>(dst);
    const uint64_t gtid = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x, gsz = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t w = gtid; w < words; w += gsz) d8[w] = s8[w];
    if (dst_ts != nullptr) for (uint64_t i = gtid; i < m; i += gsz) dst_ts[i] = src_ts[i];
    if (gtid == 0) { *n_out = m; if (n > dst_cap) atomicOr(err_flags, 2u); }
}

// ------------------------------------------------------------------------------------------------------
// synthetic stream of SURVEY.md 8d (integer arithmetic specified there; the tests check it bit for bit)
// ------------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

static __global__ void k_gen_tuple64(uint64_t seed, uint64_t start, uint32_t n, int key_mode, uint64_t nkeys,
                              const double *__restrict__ zipf_cdf, wfb_tuple64_t *__restrict__ out, uint64_t *__restrict__ ts)
{
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        const uint64_t i = start + j;
        uint64_t key;
        if (key_mode == 0) key = i % nkeys;
        else if (key_mode == 1) key = splitmix64(i) % nkeys;
        else {
            const double u = static_cast<double>(splitmix64(i ^ 0xA5A5A5A5A5A5A5A5ull) >> 11) * (1.0 / 9007199254740992.0);
            uint64_t lo = 0, hi = nkeys - 1;
            while (lo < hi) { const uint64_t mid = (lo + hi) >> 1; if (zipf_cdf[mid] > u) hi = mid; else lo = mid + 1; }
            key = lo;
        }
        const int64_t iv = static_cast<int64_t>(splitmix64(seed ^ i) & 0xFFFFull);
        const double fv = static_cast<double>(splitmix64(seed ^ ~i) >> 11) * (1.0 / 9007199254740992.0);
        uint4 *o = reinterpret_cast<uint4 *>(out + j);
        uint4 c0, c1;
        c0.x = static_cast<uint32_t>(key); c0.y = static_cast<uint32_t>(key >> 32);
        c0.z = static_cast<uint32_t>(i); c0.w = static_cast<uint32_t>(i >> 32);
        const uint64_t ivb = static_cast<uint64_t>(iv);
        const uint64_t fvb = static_cast<uint64_t>(__double_as_longlong(fv));
        c1.x = static_cast<uint32_t>(ivb); c1.y = static_cast<uint32_t>(ivb >> 32);
        c1.z = static_cast<uint32_t>(fvb); c1.w = static_cast<uint32_t>(fvb >> 32);
        o[0] = c0; o[1] = c1; o[2] = make_uint4(0, 0, 0, 0); o[3] = make_uint4(0, 0, 0, 0);
        if (ts != nullptr) ts[j] = i;
    }
}

} // namespace wfb
