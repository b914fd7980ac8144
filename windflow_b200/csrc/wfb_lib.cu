// wfb_lib.cu -- libwfb200.so: the extern "C" layer of include/wfb200.h over the kernels of wfb_kernels.cuh,
// instantiated for the built-in programs of wfb_programs.cuh. No Thrust, no unified memory, no CPU fallback:
// every compute entry point fails with WFB_E_NOGPU when there is no CUDA device.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <mutex>
#include <vector>
#include <algorithm>
#include <new>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../include/wfb200.h"
#include "wfb_kernels.cuh"
#include "wfb_launch.cuh"
#include "wfb_programs.cuh"

using namespace wfb;

#define CK(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return static_cast<int>(e__); } while (0)

namespace {

int g_num_sms = 0;

int device_ready()
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) { cudaGetLastError(); return WFB_E_NOGPU; }
    if (g_num_sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return WFB_E_NOGPU;
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    }
    return 0;
}

std::mutex &registry_mutex() { static std::mutex m; return m; } // replicas of different operators may register programs concurrently

std::vector<ProgramOps> &registry()
{
    static std::vector<ProgramOps> table; if (table.empty()) { table.reserve(4096); table = { make_ops<ProgTuple64>(), make_ops<ProgWfTest16>(), make_ops<ProgWfWin24>(), make_ops<ProgLifted32>() }; table.reserve(4096); }
    return table;
}

const ProgramOps *program(int prog)
{
    std::lock_guard<std::mutex> lock(registry_mutex());
    std::vector<ProgramOps> &t = registry();
    if (prog < 0 || prog >= static_cast<int>(t.size())) return nullptr;
    return &t[prog];
}

// ---- host -> device staging of the small per-call tables (batch descriptors, offsets): a ring of pinned buffers, so that the
// copy is a real asynchronous DMA (a cudaMemcpyAsync from pageable memory is staged by the driver and serialises with the stream) ----
struct PinnedStage {
    static constexpr int SLOTS = 8;
    unsigned char *buf[SLOTS] = {}; size_t cap[SLOTS] = {}; cudaEvent_t ev[SLOTS] = {}; bool used[SLOTS] = {}; int next = 0;
    int h2d(void *dst, const void *src, size_t bytes, cudaStream_t s)
    {
        if (bytes == 0) return 0;
        const int i = next; next = (next + 1) % SLOTS;
        if (used[i]) CK(cudaEventSynchronize(ev[i])); // the copy that used this slot SLOTS calls ago has long finished
        if (cap[i] < bytes) { if (buf[i]) cudaFreeHost(buf[i]); cap[i] = std::max<size_t>(bytes, 4096) * 2; CK(cudaMallocHost(reinterpret_cast<void **>(&buf[i]), cap[i])); }
        if (!ev[i]) CK(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
        std::memcpy(buf[i], src, bytes);
        CK(cudaMemcpyAsync(dst, buf[i], bytes, cudaMemcpyHostToDevice, s));
        CK(cudaEventRecord(ev[i], s)); used[i] = true;
        return 0;
    }
    void destroy() { for (int i = 0; i < SLOTS; i++) { if (buf[i]) cudaFreeHost(buf[i]); if (ev[i]) cudaEventDestroy(ev[i]); buf[i] = nullptr; ev[i] = nullptr; } }
};

// ---- scratch shared by the tile passes: ticket counter, epoch-tagged tile states, batch descriptors -----------
struct TileScratch {
    PinnedStage stage;
    uint64_t *tile_state = nullptr; uint32_t tile_cap = 0;
    uint32_t *ticket = nullptr; uint32_t ticket_base = 0; uint32_t epoch = 0;
    DevBatch *d_batches = nullptr; uint32_t batch_cap = 0;
    cudaStream_t last_stream = nullptr; bool used = false; cudaEvent_t ev = nullptr;

    int init()
    {
        CK(cudaMalloc(&ticket, sizeof(uint32_t)));
        CK(cudaMemset(ticket, 0, sizeof(uint32_t)));
        CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        return 0;
    }
    void destroy()
    {
        cudaFree(tile_state); cudaFree(ticket); cudaFree(d_batches);
        stage.destroy();
        if (ev) cudaEventDestroy(ev);
    }
    // scratch is shared by consecutive launches: order them when the caller hops between streams
    // (the reference launches on each batch's own stream, wf/map_gpu.hpp:399-405)
    int enter(cudaStream_t s)
    {
        if (used && s != last_stream) { CK(cudaEventRecord(ev, last_stream)); CK(cudaStreamWaitEvent(s, ev, 0)); }
        last_stream = s; used = true;
        return 0;
    }
    int ensure_tiles(uint32_t num_tiles)
    {
        if (num_tiles <= tile_cap) return 0;
        CK(cudaStreamSynchronize(last_stream));
        cudaFree(tile_state);
        tile_cap = std::max(num_tiles, 2 * tile_cap);
        CK(cudaMalloc(&tile_state, sizeof(uint64_t) * tile_cap));
        CK(cudaMemset(tile_state, 0, sizeof(uint64_t) * tile_cap)); // epoch 0 is never used by a launch
        return 0;
    }
    int ensure_batches(uint32_t nb)
    {
        if (nb <= batch_cap) return 0;
        CK(cudaStreamSynchronize(last_stream));
        cudaFree(d_batches);
        batch_cap = std::max(nb, 2 * batch_cap);
        CK(cudaMalloc(&d_batches, sizeof(DevBatch) * batch_cap));
        return 0;
    }
    void next_launch(TileArgs &a) { epoch = (epoch + 1) & 0x3fffffffu; if (epoch == 0) epoch = 1; a.epoch = epoch; a.ticket = ticket; a.ticket_base = ticket_base; a.tile_state = tile_state; }
    void launched(uint32_t num_claims, uint32_t grid) { ticket_base += num_claims + grid; } // one failing claim per CTA
};

inline uint32_t tiles_of(uint32_t n) { return (n + TILE - 1) / TILE; }

// scratch + launcher of the onesweep radix sort (one per engine / window handle)
struct RadixSorter {
    uint32_t *ctl = nullptr;        // [passes][256] histograms + [passes] tickets
    uint64_t *state = nullptr;      // [tiles][digits] look-back words
    uint64_t state_words = 0;
    uint32_t epoch = 0;
    uint64_t launches = 0;

    int ensure(uint32_t cap_elems, uint32_t min_tile, cudaStream_t s, uint32_t digits = 256)
    {
        if (!ctl) CK(cudaMalloc(&ctl, sizeof(uint32_t) * CTL_WORDS));
        const uint32_t tiles = (cap_elems + min_tile - 1) / min_tile;
        if (static_cast<uint64_t>(tiles) * digits > state_words) {
            CK(cudaStreamSynchronize(s));
            cudaFree(state);
            state_words = static_cast<uint64_t>(tiles) * digits;
            CK(cudaMalloc(&state, sizeof(uint64_t) * state_words));
            CK(cudaMemset(state, 0, sizeof(uint64_t) * state_words));
        }
        return 0;
    }
    void destroy() { cudaFree(ctl); cudaFree(state); cudaFree(wideH); cudaFree(wideC); cudaFree(wideH32); }

    // stable sort of (kA[i], i) by the low 8*passes bits; n on the device (n_ptr) or the host (n_host), cap = upper bound
    template <class K, int ITEMS>
    void launch_pass(uint32_t tiles, const K *kin, const uint32_t *vin, K *kout, uint32_t *vout, const uint32_t *n_ptr, uint32_t n_host,
                     uint32_t p, uint32_t passes, cudaStream_t s, const unsigned char *pin, unsigned char *pout, uint32_t pbytes,
                     uint32_t *seg_first, uint32_t seg_first_n, uint32_t base_shift)
    {
        k_onesweep_pass<K, ITEMS><<<tiles, OS_THREADS, 0, s>>>(kin, vin, kout, vout, n_ptr, n_host, p, passes, ctl, state, epoch,
                                                               pin, pout, pbytes, seg_first, seg_first_n, base_shift);
    }

    // clears the histograms / tickets of the next sort; call it BEFORE a producer that fills the histograms itself
    static constexpr size_t CTL_WORDS = OS_MAX_PASSES * 256 + OS_MAX_PASSES > OSW_DIGITS + 1 ? OS_MAX_PASSES * 256 + OS_MAX_PASSES : OSW_DIGITS + 1;
    static int prepare(uint32_t *ctl_buf, uint32_t passes, cudaStream_t s)
    {
        passes = std::min<uint32_t>(std::max(1u, passes), OS_MAX_PASSES);
        WFB_CK(cudaMemsetAsync(ctl_buf, 0, sizeof(uint32_t) * (passes * 256 + passes), s));
        return 0;
    }

    // wide pass (10-bit digit): ctl = [1024 counts][ticket]
    static int prepare_wide(uint32_t *ctl_buf, cudaStream_t s)
    {
        WFB_CK(cudaMemsetAsync(ctl_buf, 0, sizeof(uint32_t) * (OSW_DIGITS + 1), s));
        return 0;
    }
    // ONE stable partition pass of (kin[i], i) on the digit (key >> shift) & 1023 into (kout, vout): per-tile counts, then
    // the scatter (no chained scan). *counts = the 1024 digit counts (ready_ctl if the producer of the keys made them).
    uint16_t *wideH = nullptr; uint32_t *wideC = nullptr; uint32_t wide_tiles = 0, wide_chunks = 0;
    uint32_t *wideH32 = nullptr; uint32_t wide32_tiles = 0; // per-tile counts filled by the producer of the keys (32-bit rows)
    // rows for `cap` positions, zeroed on stream s: the producer adds its digit counts, sort_wide(..., h32_ready) consumes them
    int prepare_h32(uint32_t cap, cudaStream_t s, uint32_t **rows)
    {
        const uint32_t tiles = std::max(1u, (cap + OSW_TILE - 1) / OSW_TILE);
        if (tiles > wide32_tiles) {
            CK(cudaStreamSynchronize(s));
            cudaFree(wideH32);
            wide32_tiles = std::max(tiles, 2 * wide32_tiles);
            CK(cudaMalloc(&wideH32, sizeof(uint32_t) * OSW_DIGITS * wide32_tiles));
        }
        CK(cudaMemsetAsync(wideH32, 0, sizeof(uint32_t) * OSW_DIGITS * tiles, s));
        *rows = wideH32;
        return 0;
    }
    // tiles of the wide pass over `cap` positions and their chunks. prefix = false: about sqrt(tiles) chunks of >= 16 tiles, every
    // scatter CTA sums the rows of the earlier chunks itself; true (rows filed by the tile pass): chunks of 32 tiles whose first
    // output positions one small kernel computes (k_wide_chunk_scan), so a scatter CTA reads one chunk row and < 32 tile rows
    static void wide_geometry(uint32_t cap, bool prefix, uint32_t *tiles, uint32_t *chunk_shift, uint32_t *chunks)
    {
        *tiles = std::max(1u, (cap + OSW_TILE - 1) / OSW_TILE);
        uint32_t cs = prefix ? 5 : 4; // (prefix: 32 tiles per chunk -- the one-CTA scan over the chunks stays short, a scatter CTA adds < 32 rows)
        if (!prefix) while ((1u << (2 * cs)) < *tiles) cs++;
        *chunk_shift = cs;
        *chunks = (*tiles + (1u << cs) - 1) >> cs;
    }
    // rows of the wide partition for `cap` positions, allocated before the producer of the keys files them (TileArgs::wide_h16)
    int ensure_wide(uint32_t cap, cudaStream_t s, uint16_t **rows)
    {
        uint32_t tiles, chunk_shift, chunks;
        wide_geometry(cap, true, &tiles, &chunk_shift, &chunks);
        if (tiles > wide_tiles || chunks > wide_chunks) {
            CK(cudaStreamSynchronize(s));
            cudaFree(wideH); cudaFree(wideC);
            wide_tiles = std::max(tiles, wide_tiles); wide_chunks = std::max(chunks, wide_chunks);
            CK(cudaMalloc(&wideH, sizeof(uint16_t) * OSW_DIGITS * wide_tiles));
            CK(cudaMalloc(&wideC, sizeof(uint32_t) * OSW_DIGITS * wide_chunks));
        }
        *rows = wideH;
        return 0;
    }
    const uint16_t *rows_override = nullptr; // set by sort_wide for the duration of one call
    template <class K> static bool keys_out_ok(K *kout, uint32_t *vout) { return kout != nullptr && vout != nullptr; }
    template <class K, int RBYTES>
    void launch_wide_scatter(uint32_t tiles, const K *kin, K *kout, uint32_t *vout, const uint32_t *n_ptr, uint32_t n_host, uint32_t shift,
                             uint32_t chunk_shift, const uint32_t *c, cudaStream_t s, const unsigned char *pin, unsigned char *pout, uint32_t pbytes,
                             uint32_t skip_invalid, uint32_t region_stride, const uint32_t *h32, uint32_t cx)
    {
        k_wide_scatter<K, RBYTES><<<tiles, OSW_THREADS, 0, s>>>(kin, kout, vout, n_ptr, n_host, shift, chunk_shift, rows_override ? rows_override : wideH, wideC, c, pin, pout, pbytes, skip_invalid, region_stride, h32, cx);
    }
    // payload_in / payload_out (optional): payload_bytes-sized records that travel with the elements (multiple of 8 bytes)
    template <class K>
    int sort_wide(const K *kin, K *kout, uint32_t *vout, const uint32_t *n_ptr, uint32_t n_host, uint32_t cap, uint32_t shift,
                  cudaStream_t s, uint32_t *ready_ctl, const uint32_t **counts, const unsigned char *payload_in = nullptr,
                  unsigned char *payload_out = nullptr, uint32_t payload_bytes = 0, bool skip_invalid = false, uint32_t region_stride = 0, uint32_t few_bins = 0,
                  bool h32_ready = false, bool h16_ready = false, uint16_t *h16_rows = nullptr, bool ranked = false)
    {
        // ranked: the keys carry a rank in their upper half (TileArgs::pack_rank): k_wide_scatter_ranked
        // h16_rows: the 16-bit rows the tile pass filed when they do not live in this sorter (a pipelined handle keeps one set per segment in flight)
        if (payload_in && (payload_bytes == 0 || (payload_bytes & 7u))) return WFB_E_BADARG;
        if (!ctl) CK(cudaMalloc(&ctl, sizeof(uint32_t) * CTL_WORDS));
        const bool prefix = h16_ready && ready_ctl && !few_bins;
        rows_override = prefix ? h16_rows : nullptr;
        uint32_t tiles, chunk_shift, chunks;
        wide_geometry(cap, prefix, &tiles, &chunk_shift, &chunks);
        if (tiles > wide_tiles || chunks > wide_chunks) {
            CK(cudaStreamSynchronize(s));
            cudaFree(wideH); cudaFree(wideC);
            wide_tiles = std::max(tiles, wide_tiles); wide_chunks = std::max(chunks, wide_chunks);
            CK(cudaMalloc(&wideH, sizeof(uint16_t) * OSW_DIGITS * wide_tiles));
            CK(cudaMalloc(&wideC, sizeof(uint32_t) * OSW_DIGITS * wide_chunks));
        }
        uint32_t *c = ready_ctl ? ready_ctl : ctl;
        if (!ready_ctl) { int rc = prepare_wide(c, s); if (rc) return rc; }
        const uint32_t *h32 = nullptr;
        if (h16_ready && ready_ctl && few_bins) { // a few bins (destinations), rows filed by the tile pass: chunk sums + the global counts (c was cleared by the caller)
            k_wide_chunk_sums16<<<chunks, OSW_THREADS, 0, s>>>(h16_rows ? h16_rows : wideH, tiles, chunk_shift, wideC, c);
        } else if (prefix) { // the tile pass filed the 16-bit rows (wideH): chunk sums, then first output position of every (chunk, digit) + digit counts
            k_wide_chunk_sums16<<<chunks, OSW_THREADS, 0, s>>>(h16_rows ? h16_rows : wideH, tiles, chunk_shift, wideC);
            k_wide_chunk_scan<<<1, OSW_DIGITS, 0, s>>>(wideC, chunks, c);
            launches++;
        } else if (h32_ready && ready_ctl && !few_bins) { // the producer of the keys counted the digits per tile: only the chunk sums are missing
            k_wide_chunk_sums<<<chunks, OSW_THREADS, 0, s>>>(wideH32, tiles, chunk_shift, wideC);
            h32 = wideH32;
        } else {
            CK(cudaMemsetAsync(wideC, 0, sizeof(uint32_t) * OSW_DIGITS * chunks, s));
            k_wide_tile_hist<K><<<tiles, OSW_THREADS, 0, s>>>(kin, n_ptr, n_host, shift, chunk_shift, wideH, wideC, ready_ctl ? nullptr : c, skip_invalid ? 1u : 0u);
        }
        if (region_stride && few_bins && payload_in && few_bins <= 32 && (payload_bytes == 16 || payload_bytes == 24 || payload_bytes == 32 || payload_bytes == 64)) {
            // a few fixed-capacity regions (destination GPUs): ranks from ballots, records leave in runs
            const uint32_t *d32 = reinterpret_cast<const uint32_t *>(kin);
            static_assert(sizeof(K) == 4 || sizeof(K) == 8, "key width");
            if (sizeof(K) == 4) {
                switch (payload_bytes) {
                    case 16: k_shard_scatter<16><<<tiles, OSW_THREADS, 0, s>>>(d32, n_host, few_bins, chunk_shift, wideH, wideC, payload_in, payload_out, region_stride); break;
                    case 24: k_shard_scatter<24><<<tiles, OSW_THREADS, 0, s>>>(d32, n_host, few_bins, chunk_shift, wideH, wideC, payload_in, payload_out, region_stride); break;
                    case 32: k_shard_scatter<32><<<tiles, OSW_THREADS, 0, s>>>(d32, n_host, few_bins, chunk_shift, wideH, wideC, payload_in, payload_out, region_stride); break;
                    default: k_shard_scatter<64><<<tiles, OSW_THREADS, 0, s>>>(d32, n_host, few_bins, chunk_shift, wideH, wideC, payload_in, payload_out, region_stride); break;
                }
                CK(cudaGetLastError());
                launches += 2;
                *counts = c;
                return 0;
            }
        }
        if (prefix && ranked && sizeof(K) == 4 && !payload_in && keys_out_ok(kout, vout)) {
            k_wide_scatter_ranked<0><<<tiles, OSW_THREADS, 0, s>>>(reinterpret_cast<const uint32_t *>(kin), reinterpret_cast<uint32_t *>(kout), vout, n_host, shift, chunk_shift,
                                                                   h16_rows ? h16_rows : wideH, wideC, nullptr, nullptr);
            CK(cudaGetLastError());
            launches += 2;
            *counts = c;
            return 0;
        }
        if (prefix && ranked && sizeof(K) == 4 && payload_in && payload_out && kout && !region_stride) { // the records travel with their slots (bucketed exchange)
#define WFB_WSR(RB_) k_wide_scatter_ranked<RB_><<<tiles, OSW_THREADS, 0, s>>>(reinterpret_cast<const uint32_t *>(kin), reinterpret_cast<uint32_t *>(kout), nullptr, n_host, shift, \
                                                                             chunk_shift, h16_rows ? h16_rows : wideH, wideC, payload_in, payload_out)
            switch (payload_bytes) {
                case 16: WFB_WSR(16); break; case 24: WFB_WSR(24); break; case 32: WFB_WSR(32); break; case 48: WFB_WSR(48); break; case 64: WFB_WSR(64); break;
                default: return WFB_E_UNSUPPORTED;
            }
#undef WFB_WSR
            CK(cudaGetLastError());
            launches += 2;
            *counts = c;
            return 0;
        }
#define WFB_WS(RB_) launch_wide_scatter<K, RB_>(tiles, kin, kout, vout, n_ptr, n_host, shift, chunk_shift, c, s, payload_in, payload_out, payload_bytes, skip_invalid ? 1u : 0u, region_stride, h32, prefix ? 1u : 0u)
        if (!payload_in) WFB_WS(0);
        else switch (payload_bytes) {
            case 8: WFB_WS(8); break;   case 16: WFB_WS(16); break; case 24: WFB_WS(24); break; case 32: WFB_WS(32); break;
            case 48: WFB_WS(48); break; case 64: WFB_WS(64); break; default: WFB_WS(-1); break;
        }
#undef WFB_WS
        CK(cudaGetLastError());
        launches += 2;
        *counts = c;
        return 0;
    }

    template <class K>
    int sort(K *kA, K *kB, uint32_t *vA, uint32_t *vB, const uint32_t *n_ptr, uint32_t n_host, uint32_t cap, uint32_t passes,
             cudaStream_t s, const K **skeys, const uint32_t **svals,
             const unsigned char *payload_in = nullptr, unsigned char *payload_out = nullptr, uint32_t payload_bytes = 0,
             uint32_t *ready_ctl = nullptr, uint32_t *seg_first = nullptr, uint32_t seg_first_n = 0, uint32_t base_shift = 0)
    {
        // ready_ctl != nullptr: histograms already accumulated there by the producer of the keys (after prepare())
        uint32_t *const own_ctl = ctl;
        const bool hist_ready = ready_ctl != nullptr;
        if (hist_ready) ctl = ready_ctl;
        struct Restore { uint32_t *&c; uint32_t *v; ~Restore() { c = v; } } restore{ctl, own_ctl};
        static int items = 0; // elements per thread of a pass (tuning knob: WFB_OS_ITEMS = 4, 8 or 16)
        if (items == 0) {
            const char *e = std::getenv("WFB_OS_ITEMS");
            items = e ? std::atoi(e) : 8;
            if (items != 4 && items != 8 && items != 16) items = 8;
            if (items > OsCfg<K>::MAX_ITEMS) items = OsCfg<K>::MAX_ITEMS;
        }
        const uint32_t TE = OS_THREADS * static_cast<uint32_t>(items);
        int rc = ensure(cap, OS_THREADS * 4, s, 256); if (rc) return rc;
        passes = std::min<uint32_t>(std::max(1u, passes), OS_MAX_PASSES);
        const uint32_t tiles = std::max(1u, (cap + TE - 1) / TE);
        if (!hist_ready) {
            CK(cudaMemsetAsync(ctl, 0, sizeof(uint32_t) * (passes * 256 + passes), s));
            k_radix_ghist<K><<<std::min(tiles, static_cast<uint32_t>(g_num_sms) * 4u), 256, 0, s>>>(kA, n_ptr, n_host, passes, ctl, base_shift);
            launches++;
        }
        const K *kin = kA; const uint32_t *vin = nullptr;
        K *kout = kB; uint32_t *vout = vB;
        for (uint32_t p = 0; p < passes; p++) {
            epoch = (epoch + 1) & 0x3fffffffu; if (epoch == 0) epoch = 1;
            const bool last = (p + 1 == passes);
            const unsigned char *pin = last ? payload_in : nullptr;
            unsigned char *pout = last ? payload_out : nullptr;
            uint32_t *sf = last ? seg_first : nullptr;
            if (items == 4) launch_pass<K, 4>(tiles, kin, vin, kout, vout, n_ptr, n_host, p, passes, s, pin, pout, payload_bytes, sf, seg_first_n, base_shift);
            else if (items == 8 || OsCfg<K>::MAX_ITEMS < 16) launch_pass<K, 8>(tiles, kin, vin, kout, vout, n_ptr, n_host, p, passes, s, pin, pout, payload_bytes, sf, seg_first_n, base_shift);
            else launch_pass<K, (OsCfg<K>::MAX_ITEMS >= 16 ? 16 : 8)>(tiles, kin, vin, kout, vout, n_ptr, n_host, p, passes, s, pin, pout, payload_bytes, sf, seg_first_n, base_shift);
            kin = kout; vin = vout;
            if (kout == kB) { kout = kA; vout = vA; } else { kout = kB; vout = vB; }
        }
        CK(cudaGetLastError());
        launches += passes;
        *skeys = kin; *svals = vin;
        return 0;
    }
};

} // namespace

struct wfb_engine {
    int prog = 0;
    const ProgramOps *ops = nullptr;
    std::vector<unsigned char> params; // the program's params_t used by key / reduce (zeros unless wfb_engine_set_params)
    const void *pp() const { return params.empty() ? nullptr : params.data(); }
    TileScratch ts;
    uint64_t launches = 0;
    // scratch of the per-batch keyed operators (sort buffers), grown on demand
    uint32_t key_bits = 64;
    uint32_t cap = 0;
    uint64_t *keysA = nullptr, *keysB = nullptr;
    uint32_t *idxA = nullptr, *idxB = nullptr, *destA = nullptr, *destB = nullptr;
    uint32_t *head = nullptr, *seg_begin = nullptr;
    // wfb_shard_lift: lifted records / destinations of one segment, tile t owns positions [t*TILE, +TILE)
    unsigned char *sh_lifted = nullptr; uint32_t *sh_dest = nullptr, *sh_ctl = nullptr; uint64_t sh_cap = 0;
    // wfb_reduce_by_key_batches: element offset of every batch, first segment of every batch, segment total
    uint32_t *rb_off = nullptr, *rb_first = nullptr, *rb_total = nullptr; uint32_t rb_cap = 0;
    uint32_t *rb_long = nullptr; uint32_t rb_long_cap = 0; // segments folded by a warp; rb_total[1] = their number
    RadixSorter sorter;

    int ensure_sort(uint32_t n, cudaStream_t s)
    {
        if (n <= cap) return 0;
        CK(cudaStreamSynchronize(s));
        cudaFree(keysA); cudaFree(keysB); cudaFree(idxA); cudaFree(idxB); cudaFree(destA); cudaFree(destB);
        cudaFree(head); cudaFree(seg_begin);
        cap = std::max(n, 2 * cap);
        CK(cudaMalloc(&keysA, sizeof(uint64_t) * cap)); CK(cudaMalloc(&keysB, sizeof(uint64_t) * cap));
        CK(cudaMalloc(&idxA, sizeof(uint32_t) * cap)); CK(cudaMalloc(&idxB, sizeof(uint32_t) * cap));
        CK(cudaMalloc(&destA, sizeof(uint32_t) * cap)); CK(cudaMalloc(&destB, sizeof(uint32_t) * cap));
        CK(cudaMalloc(&head, sizeof(uint32_t) * cap)); CK(cudaMalloc(&seg_begin, sizeof(uint32_t) * (static_cast<size_t>(cap) + 1)));
        return 0;
    }
    // stable LSD radix sort of (keysA[i], i) by key over `key_bits` bits; returns the buffers holding the result
    int sort64(uint32_t n, cudaStream_t s, const uint64_t **skeys, const uint32_t **sidx)
    {
        const uint64_t before = sorter.launches;
        int rc = sorter.sort<uint64_t>(keysA, keysB, idxA, idxB, nullptr, n, n, (key_bits + 7) / 8, s, skeys, sidx);
        launches += sorter.launches - before;
        return rc;
    }
    void free_sort()
    {
        cudaFree(keysA); cudaFree(keysB); cudaFree(idxA); cudaFree(idxB); cudaFree(destA); cudaFree(destB);
        cudaFree(head); cudaFree(seg_begin);
        cudaFree(sh_lifted); cudaFree(sh_dest); cudaFree(sh_ctl);
        cudaFree(rb_off); cudaFree(rb_first); cudaFree(rb_total); cudaFree(rb_long);
        sorter.destroy();
    }
};

// per-segment scratch of one Ffat_Windows_GPU; two sets when the handle is pipelined (the ingest pass of segment k+1
// overlaps sort + update of segment k)
struct SegScratch {
    uint32_t cap = 0;                     // capacity in records
    unsigned char *lifted = nullptr, *lifted_sorted = nullptr;
    uint32_t *slotsA = nullptr, *slotsB = nullptr, *posA = nullptr, *posB = nullptr;
    uint32_t *batch_off = nullptr; uint32_t batch_off_cap = 0;
    DevBatch *d_batches = nullptr; uint32_t batch_cap = 0;
    uint32_t *n_total = nullptr;
    bool sparse = false;          // this segment was ingested without global compaction (positions = tuple indices)
    bool h32_ready = false;       // the streaming pass filed the per-tile digit counts of the wide partition (sorter.wideH32)
    bool h16_ready = false;       // the tile pass filed them per wide tile as 16-bit rows (sorter.wideH), claiming 16 tiles per ticket
    bool ranked = false;          // ... and packed a rank with every slot (TileArgs::pack_rank)
    uint16_t *h16 = nullptr; uint32_t h16_tiles = 0; // pipelined handles: this segment's own rows (the next segment's tile pass files its rows while
                                                     // this segment's partition still reads these)
    const unsigned char *lifted_src = nullptr; // records of this segment: `lifted`, or the caller's buffer (in-place ingest)
    uint32_t *seg_cnt = nullptr;          // per-slot item counts of the segment (max_keys)
    Trigger *trig = nullptr; uint32_t *n_trig = nullptr; uint32_t trig_cap = 0;
    uint32_t *n_heavy = nullptr;
    uint32_t *sort_ctl = nullptr;         // digit histograms + tickets of this segment's slot sort
    // pipelined mode: results of the segment wait here until the next call / flush delivers them
    unsigned char *res = nullptr; uint64_t *res_ts = nullptr; uint32_t *res_n = nullptr; uint32_t res_cap = 0;
    cudaEvent_t ev_ingest = nullptr, ev_done = nullptr;
    bool pending = false, hist_ready = false;
    uint32_t nbatches = 0, total = 0;

    void destroy()
    {
        cudaFree(lifted); cudaFree(lifted_sorted); cudaFree(slotsA); cudaFree(slotsB); cudaFree(posA); cudaFree(posB);
        cudaFree(batch_off); cudaFree(d_batches); cudaFree(n_total); cudaFree(seg_cnt); cudaFree(trig); cudaFree(sort_ctl);
        cudaFree(res); cudaFree(res_ts); cudaFree(h16); // n_trig and res_n live inside the n_total allocation
        if (ev_ingest) cudaEventDestroy(ev_ingest);
        if (ev_done) cudaEventDestroy(ev_done);
    }
};

struct wfb_ffat {
    int prog = 0;
    const ProgramOps *ops = nullptr;
    std::vector<unsigned char> params; // the program's params_t used by key / lift / comb (zeros unless wfb_ffat_set_params)
    const void *pp() const { return params.empty() ? nullptr : params.data(); }
    FfatDev ff{};
    TileScratch ts;
    uint32_t sort_passes = 1;
    SegScratch seg[2];
    RadixSorter sorter;
    bool pipelined = false;
    uint64_t call_no = 0;
    cudaStream_t s2 = nullptr;            // pipelined mode: sort + update + window queries run here
    uint64_t launches = 0;
    size_t state_bytes = 0;
    int win_type = 0;
    // time-based windows (win_type 1): this handle is the front end (key table, rings of pending panes); the popped panes go to
    // `cb`, a count-based handle over the lifted program with window / slide in panes
    TbDev tb{};
    wfb_ffat *cb = nullptr;
    uint64_t tb_lateness = 0;
    uint32_t tb_cap = 0, tb_pop_cap = 0;            // scratch capacities (tuples per batch, popped panes)
    uint64_t *tb_kA = nullptr, *tb_kB = nullptr; uint32_t *tb_iA = nullptr, *tb_iB = nullptr;
    unsigned char *tb_lifted = nullptr, *tb_part = nullptr, *tb_popped = nullptr;
    uint32_t *tb_popped_slots = nullptr;
    bool shares_slot_key = false;                   // back end of a time-based handle: ff.slot_key / ff.n_slots belong to the front end
    uint32_t *own_n_slots = nullptr;                // the allocation behind ff.n_slots / ff.err_flags of this handle
    uint32_t *tb_head = nullptr, *tb_seg = nullptr, *tb_misc = nullptr; // misc: [0] n_segs [1] first_seg dummy [2] n_present [3] popped total [4] ignored [5] ring capacity needed
    uint32_t *mg_scratch = nullptr; // (internal, ffat_process_prebucketed) per-(bucket, sub-bucket, source) counts, their scan, run starts
    bool append_results = false;  // (internal, wfb_mg_flush) the next call's results follow the ones already in the output buffer
    bool buckets = true;          // one wide radix pass + per-bucket CTAs (<= 65536 keys); WFB_UPDATE=lanes selects the
                                  // full sort + thread-per-key update instead
    uint32_t bucket_shift = 0;    // the wide pass partitions on (slot >> bucket_shift) & 1023
    bool bucket_move = false;     // WFB_BUCKET_MOVE=1: the wide pass also moves the lifted records into their buckets
    bool l2_hints = true;         // WFB_L2_HINTS=0: no eviction-priority hints on the ingest pass
    bool fuse_tile_hist = false;  // WFB_FUSE_TILE_HIST=1: the tile pass also files the per-tile digit counts of the wide partition (one more
                                  // global RED per survivor: measured +17 us on the tile pass against -8 us on the partition, so off)
    bool stream_update = false;   // WFB_UPDATE=stream: k_ffat_update_stream (lane = key, per-key queues) instead of k_ffat_update_buckets
    bool rank_scatter = true;     // with tile_h16: the tile pass also packs a rank with every slot and the partition places the pairs by it (k_wide_scatter_ranked); WFB_RANK_SCATTER=0: off
    bool tile_h16 = true;         // the tile pass claims whole wide tiles and files their digit counts itself (no k_wide_tile_hist); WFB_TILE_H16=0: off
    bool inplace_kernel = true;   // WFB_INPLACE_KERNEL=0: the in-place case also goes through the tile pass
    bool inplace_ok = true;       // WFB_INPLACE=0: always copy the records of a pass-through program
    bool sparse_ingest = true;    // WFB_SPARSE=0: the bucket path also compacts the survivors over the whole segment
    uint32_t ingest_ctas_per_sm = 0; // 0: as many as fit; pipelined handles leave room for the concurrent sort/update kernels
    bool move_payload = false;    // tuning knob WFB_SORT_PAYLOAD=1: the last sort pass also moves the lifted records
    // optional per-phase timing (wfb_ffat_timing)
    bool timing = false;
    std::vector<cudaEvent_t> tev; // 4 events per recorded call
    uint32_t tev_used = 0;        // recorded calls since the last query
    static constexpr uint32_t TEV_MAX = 512;
    void mark(int which, cudaStream_t s)
    {
        if (!timing || tev_used >= TEV_MAX) return;
        cudaEventRecord(tev[tev_used * 4 + which], s);
    }
};

extern "C" {

int wfb_abi_version(void) { return WFB_ABI_VERSION; }

const char *wfb_error_string(int code)
{
    switch (code) {
    case 0: return "success";
    case WFB_E_BADARG: return "wfb: bad argument";
    case WFB_E_NOPROG: return "wfb: unknown program id";
    case WFB_E_CAPACITY: return "wfb: capacity exceeded (keys or results)";
    case WFB_E_NOGPU: return "wfb: no CUDA device (libwfb200 has no CPU fallback)";
    case WFB_E_UNSUPPORTED: return "wfb: not supported";
    }
    if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
    return "wfb: unknown error";
}

int wfb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int wfb_program_register(const void *ops, size_t ops_bytes)
{
    if (!ops || ops_bytes != sizeof(ProgramOps)) return WFB_E_BADARG;
    std::lock_guard<std::mutex> lock(registry_mutex());
    std::vector<ProgramOps> &t = registry();
    if (t.size() >= 4096) return WFB_E_CAPACITY;
    t.reserve(4096); // handles keep pointers into the table: never reallocate it
    t.push_back(*static_cast<const ProgramOps *>(ops));
    return static_cast<int>(t.size()) - 1;
}

int wfb_program_info(int prog, wfb_program_info_t *info)
{
    const ProgramOps *o = program(prog);
    if (!o) return WFB_E_NOPROG;
    if (!info) return WFB_E_BADARG;
    info->tuple_bytes = o->tuple_bytes; info->result_bytes = o->result_bytes; info->key_bytes = 8; info->reserved = 0;
    return 0;
}

// ---- engine -------------------------------------------------------------------------------------------------
int wfb_engine_create(wfb_engine_t **e, int prog)
{
    if (!e) return WFB_E_BADARG;
    const ProgramOps *o = program(prog);
    if (!o) return WFB_E_NOPROG;
    int rc = device_ready(); if (rc) return rc;
    wfb_engine *g = new (std::nothrow) wfb_engine();
    if (!g) return WFB_E_BADARG;
    g->prog = prog; g->ops = o;
    rc = g->ts.init(); if (rc) { delete g; return rc; }
    *e = g;
    return 0;
}

int wfb_engine_destroy(wfb_engine_t *e)
{
    if (!e) return 0;
    cudaDeviceSynchronize();
    e->ts.destroy();
    e->free_sort();
    delete e;
    return 0;
}

uint64_t wfb_engine_launches(const wfb_engine_t *e) { return e ? e->launches : 0; }

int wfb_engine_set_params(wfb_engine_t *e, const void *params, size_t bytes)
{
    if (!e || !params || bytes != e->ops->params_bytes) return WFB_E_BADARG;
    e->params.assign(static_cast<const unsigned char *>(params), static_cast<const unsigned char *>(params) + bytes);
    return 0;
}

static int run_single(wfb_engine_t *e, int mode, const wfb_functors_t *f, const DevBatch &b, cudaStream_t s)
{
    const uint32_t num_tiles = tiles_of(b.n);
    int rc = e->ts.enter(s); if (rc) return rc;
    rc = e->ts.ensure_tiles(num_tiles); if (rc) return rc;
    TileArgs a; std::memset(&a, 0, sizeof(a));
    a.batches = nullptr; a.one = b; a.nbatches = 1; a.num_tiles = num_tiles;
    e->ts.next_launch(a);
    uint32_t grid = 0;
    const uint64_t sb = reinterpret_cast<uint64_t>(b.tuples);
    rc = e->ops->tile_pass(mode, a, f, num_tiles, s, &grid, sb, sb + static_cast<uint64_t>(b.n) * e->ops->tuple_bytes); if (rc) return rc;
    e->ts.launched(num_tiles, grid);
    e->launches++;
    return 0;
}

int wfb_map(wfb_engine_t *e, const wfb_functors_t *f, void *tuples, uint32_t n, void *stream)
{
    if (!e || !f || (!tuples && n)) return WFB_E_BADARG;
    if (n == 0) return 0;
    DevBatch b; std::memset(&b, 0, sizeof(b));
    b.tuples = static_cast<const unsigned char *>(tuples); b.out = static_cast<unsigned char *>(tuples); b.n = n;
    return run_single(e, MODE_MAP, f, b, static_cast<cudaStream_t>(stream));
}

int wfb_map_filter(wfb_engine_t *e, const wfb_functors_t *f, const void *tuples_in, const uint64_t *ts_in, uint32_t n,
                   void *tuples_out, uint64_t *ts_out, uint32_t *n_out_dev, void *stream)
{
    if (!e || !f || !n_out_dev || (n && (!tuples_in || !tuples_out))) return WFB_E_BADARG;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (n == 0) { CK(cudaMemsetAsync(n_out_dev, 0, sizeof(uint32_t), s)); return 0; }
    DevBatch b; std::memset(&b, 0, sizeof(b));
    b.tuples = static_cast<const unsigned char *>(tuples_in); b.ts = ts_in;
    b.out = static_cast<unsigned char *>(tuples_out); b.ts_out = ts_in ? ts_out : nullptr; b.n_out = n_out_dev; b.n = n;
    return run_single(e, MODE_FILTER, f, b, s);
}

int wfb_map_filter_batches(wfb_engine_t *e, const wfb_functors_t *f, const wfb_batch_t *in_h, const wfb_batch_t *out_h, uint32_t nbatches,
                           uint32_t *n_out_dev, void *stream)
{
    if (!e || !f || !n_out_dev || (nbatches && (!in_h || !out_h))) return WFB_E_BADARG;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (nbatches == 0) return 0;
    int rc = e->ts.enter(s); if (rc) return rc;
    CK(cudaMemsetAsync(n_out_dev, 0, sizeof(uint32_t) * nbatches, s)); // empty batches keep 0
    std::vector<DevBatch> hb; hb.reserve(nbatches);
    uint32_t tiles = 0; uint64_t span_begin = ~0ull, span_end = 0;
    for (uint32_t i = 0; i < nbatches; i++) {
        if (in_h[i].n == 0) continue;
        if (!in_h[i].tuples || !out_h[i].tuples) return WFB_E_BADARG;
        DevBatch b; std::memset(&b, 0, sizeof(b));
        b.tuples = static_cast<const unsigned char *>(in_h[i].tuples); b.ts = in_h[i].ts;
        b.out = static_cast<unsigned char *>(const_cast<void *>(out_h[i].tuples));
        b.ts_out = in_h[i].ts ? const_cast<uint64_t *>(out_h[i].ts) : nullptr;
        b.n_out = n_out_dev + i; b.n = in_h[i].n; b.tile_begin = tiles;
        tiles += tiles_of(b.n);
        const uint64_t p0 = reinterpret_cast<uint64_t>(b.tuples);
        span_begin = std::min(span_begin, p0); span_end = std::max(span_end, p0 + static_cast<uint64_t>(b.n) * e->ops->tuple_bytes);
        hb.push_back(b);
    }
    if (hb.empty()) return 0;
    rc = e->ts.ensure_tiles(tiles); if (rc) return rc;
    rc = e->ts.ensure_batches(static_cast<uint32_t>(hb.size())); if (rc) return rc;
    { int rc_ = e->ts.stage.h2d(e->ts.d_batches, hb.data(), sizeof(DevBatch) * hb.size(), s); if (rc_) return rc_; }
    TileArgs a; std::memset(&a, 0, sizeof(a));
    a.batches = e->ts.d_batches; a.nbatches = static_cast<uint32_t>(hb.size()); a.num_tiles = tiles; a.l2_hints = 1;
    e->ts.next_launch(a);
    uint32_t grid = 0;
    rc = e->ops->tile_pass(MODE_FILTER, a, f, tiles, s, &grid, span_begin, span_end); if (rc) return rc;
    e->ts.launched(tiles, grid);
    e->launches++;
    return 0;
}

int wfb_engine_set_key_bits(wfb_engine_t *e, uint32_t bits)
{
    if (!e || bits == 0 || bits > 64) return WFB_E_BADARG;
    e->key_bits = bits;
    return 0;
}

// sort the batch's (key, index) pairs and derive the key segments; shared by reduce_by_key and keyby_group
static int keyed_prepare(wfb_engine_t *e, const void *tuples, uint32_t n, int32_t *map_idxs, int32_t *start_idxs, uint64_t *dist_keys,
                         uint32_t *n_keys_dev, cudaStream_t s, const uint32_t **sidx_out)
{
    int rc = e->ts.enter(s); if (rc) return rc;
    rc = e->ensure_sort(n, s); if (rc) return rc;
    rc = e->ops->extract_keys(static_cast<const unsigned char *>(tuples), n, e->keysA, nullptr, 1, s, e->pp()); if (rc) return rc;
    e->launches++;
    const uint64_t *skeys; const uint32_t *sidx;
    rc = e->sort64(n, s, &skeys, &sidx); if (rc) return rc;
    const uint32_t g = grid_for(n, 256);
    k_seg_heads<<<g, 256, 0, s>>>(skeys, sidx, n, e->head, map_idxs);
    k_scan_u32<<<1, 1024, 0, s>>>(e->head, e->head, n, nullptr);
    k_seg_finish<<<g, 256, 0, s>>>(skeys, sidx, e->head, n, start_idxs, dist_keys, e->seg_begin, n_keys_dev);
    CK(cudaGetLastError());
    e->launches += 3;
    *sidx_out = sidx;
    return 0;
}

int wfb_reduce_by_key(wfb_engine_t *e, const void *tuples, const uint64_t *ts, uint32_t n,
                      void *out_tuples, uint64_t *out_ts, uint32_t *n_out_dev, void *stream)
{
    if (!e || !n_out_dev || (n && (!tuples || !out_tuples))) return WFB_E_BADARG;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (n == 0) { CK(cudaMemsetAsync(n_out_dev, 0, sizeof(uint32_t), s)); return 0; }
    const uint32_t *sidx;
    int rc = keyed_prepare(e, tuples, n, nullptr, nullptr, nullptr, n_out_dev, s, &sidx); if (rc) return rc;
    rc = e->ops->reduce_segments(static_cast<const unsigned char *>(tuples), ts, sidx, e->seg_begin, n_out_dev,
                                 static_cast<unsigned char *>(out_tuples), ts ? out_ts : nullptr, n, s, e->pp());
    if (rc) return rc;
    e->launches++;
    return 0;
}

int wfb_reduce_by_key_batches(wfb_engine_t *e, const wfb_batch_t *in_h, const wfb_batch_t *out_h, uint32_t nbatches, uint32_t *n_out_dev,
                              void *stream)
{
    if (!e || !n_out_dev || (nbatches && (!in_h || !out_h))) return WFB_E_BADARG;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (nbatches == 0) return 0;
    uint32_t bbits = 0; while ((1ull << bbits) < nbatches) bbits++;
    if (e->key_bits + bbits > 64) return WFB_E_BADARG;
    int rc = e->ts.enter(s); if (rc) return rc;
    CK(cudaMemsetAsync(n_out_dev, 0, sizeof(uint32_t) * nbatches, s));
    std::vector<DevBatch> hb(nbatches);
    std::vector<uint32_t> boff(nbatches + 1);
    uint64_t total = 0;
    for (uint32_t i = 0; i < nbatches; i++) { // (empty batches keep their index: the composite key carries it)
        if (in_h[i].n && (!in_h[i].tuples || !out_h[i].tuples)) return WFB_E_BADARG;
        DevBatch b; std::memset(&b, 0, sizeof(b));
        b.tuples = static_cast<const unsigned char *>(in_h[i].tuples); b.ts = in_h[i].ts;
        b.out = static_cast<unsigned char *>(const_cast<void *>(out_h[i].tuples));
        b.ts_out = in_h[i].ts ? const_cast<uint64_t *>(out_h[i].ts) : nullptr;
        b.n_out = n_out_dev + i; b.n = in_h[i].n;
        hb[i] = b; boff[i] = static_cast<uint32_t>(total); total += b.n;
    }
    boff[nbatches] = static_cast<uint32_t>(total);
    if (total == 0) return 0;
    if (total > 0x7fffffffull) return WFB_E_BADARG;
    const uint32_t n = static_cast<uint32_t>(total);
    rc = e->ensure_sort(n, s); if (rc) return rc;
    rc = e->ts.ensure_batches(nbatches); if (rc) return rc;
    if (nbatches > e->rb_cap) {
        CK(cudaStreamSynchronize(s));
        cudaFree(e->rb_off); cudaFree(e->rb_first);
        e->rb_cap = std::max(nbatches, 2 * e->rb_cap);
        CK(cudaMalloc(&e->rb_off, sizeof(uint32_t) * (static_cast<size_t>(e->rb_cap) + 1)));
        CK(cudaMalloc(&e->rb_first, sizeof(uint32_t) * e->rb_cap));
        if (!e->rb_total) CK(cudaMalloc(&e->rb_total, sizeof(uint32_t) * 2));
    }
    { int rc_ = e->ts.stage.h2d(e->ts.d_batches, hb.data(), sizeof(DevBatch) * nbatches, s); if (rc_) return rc_; }
    { int rc_ = e->ts.stage.h2d(e->rb_off, boff.data(), sizeof(uint32_t) * (nbatches + 1), s); if (rc_) return rc_; }
    CK(cudaMemsetAsync(e->rb_first, 0xff, sizeof(uint32_t) * nbatches, s));
    CK(cudaMemsetAsync(e->rb_total, 0, sizeof(uint32_t) * 2, s));
    if (n / RB_LONG + 1 > e->rb_long_cap) {
        CK(cudaStreamSynchronize(s));
        cudaFree(e->rb_long);
        e->rb_long_cap = std::max(n / RB_LONG + 1, 2 * e->rb_long_cap);
        CK(cudaMalloc(&e->rb_long, sizeof(uint32_t) * e->rb_long_cap));
    }
    const uint32_t kb = nbatches == 1 ? 64u : e->key_bits; // a single batch needs no composite key
    rc = e->ops->extract_keys_batches(e->ts.d_batches, e->rb_off, nbatches, n, kb, e->keysA, s, e->pp()); if (rc) return rc;
    const uint64_t *skeys; const uint32_t *sidx;
    const uint64_t before = e->sorter.launches;
    const uint32_t sort_bits = nbatches == 1 ? e->key_bits : e->key_bits + bbits;
    rc = e->sorter.sort<uint64_t>(e->keysA, e->keysB, e->idxA, e->idxB, nullptr, n, n, (sort_bits + 7) / 8, s, &skeys, &sidx); if (rc) return rc;
    const uint32_t tiles = (n + SEGT - 1) / SEGT;
    k_head_tile_counts<<<tiles, 256, 0, s>>>(skeys, n, e->head);
    k_scan_u32<<<1, 1024, 0, s>>>(e->head, e->head, tiles, nullptr);
    k_seg_finish_batches<<<tiles, 256, 0, s>>>(skeys, n, kb, e->head, e->seg_begin, e->rb_first, e->rb_total);
    k_batch_seg_counts<<<1, 32, 0, s>>>(e->rb_first, nbatches, e->rb_total, e->ts.d_batches);
    CK(cudaGetLastError());
    rc = e->ops->reduce_segments_batches(e->ts.d_batches, e->rb_off, skeys, sidx, e->seg_begin, e->rb_first, e->rb_total, kb, n, e->rb_long,
                                         e->rb_total + 1, s, e->pp());
    if (rc) return rc;
    e->launches += 7 + (e->sorter.launches - before);
    return 0;
}

int wfb_reduce_all(wfb_engine_t *e, const void *tuples, const uint64_t *ts, uint32_t n, void *out_tuple, uint64_t *out_ts, void *stream)
{
    if (!e || !out_tuple || (n && !tuples)) return WFB_E_BADARG;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int rc = e->ts.enter(s); if (rc) return rc;
    rc = e->ops->reduce_all(static_cast<const unsigned char *>(tuples), ts, n, static_cast<unsigned char *>(out_tuple), out_ts, s, e->pp());
    if (rc) return rc;
    e->launches++;
    return 0;
}

int wfb_keyby_group(wfb_engine_t *e, const void *tuples, uint32_t n, int32_t *start_idxs, int32_t *map_idxs, uint64_t *dist_keys,
                    uint32_t *n_keys_dev, void *stream)
{
    if (!e || !n_keys_dev || (n && (!tuples || !start_idxs || !map_idxs || !dist_keys))) return WFB_E_BADARG;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (n == 0) { CK(cudaMemsetAsync(n_keys_dev, 0, sizeof(uint32_t), s)); return 0; }
    const uint32_t *sidx;
    return keyed_prepare(e, tuples, n, map_idxs, start_idxs, dist_keys, n_keys_dev, s, &sidx);
}

int wfb_shard_by_key(wfb_engine_t *e, const void *tuples, const uint64_t *ts, uint32_t n, uint32_t num_shards,
                     void *out_tuples, uint64_t *out_ts, uint32_t *seg_off_dev, void *stream)
{
    if (!e || !seg_off_dev || num_shards == 0 || num_shards > 256 || (n && (!tuples || !out_tuples))) return WFB_E_BADARG;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (n == 0) { CK(cudaMemsetAsync(seg_off_dev, 0, sizeof(uint32_t) * (num_shards + 1), s)); return 0; }
    int rc = e->ts.enter(s); if (rc) return rc;
    rc = e->ensure_sort(n, s); if (rc) return rc;
    rc = e->ops->extract_keys(static_cast<const unsigned char *>(tuples), n, nullptr, e->destA, num_shards, s, e->pp()); if (rc) return rc;
    const uint32_t *sdest, *perm;
    const uint64_t before = e->sorter.launches;
    rc = e->sorter.sort<uint32_t>(e->destA, e->destB, e->idxA, e->idxB, nullptr, n, n, 1, s, &sdest, &perm); if (rc) return rc;
    k_shard_offsets<<<1, 288, 0, s>>>(sdest, n, num_shards, seg_off_dev);
    CK(cudaGetLastError());
    rc = e->ops->gather(static_cast<const unsigned char *>(tuples), ts, perm, n, static_cast<unsigned char *>(out_tuples),
                        ts ? out_ts : nullptr, s);
    if (rc) return rc;
    e->launches += 3 + (e->sorter.launches - before);
    return 0;
}

// bucketed mode (wfb_mg_*): shard_slots != 0 -- the partition is the full 1024-bin one on the destination-major virtual slot
// (dest * shard_slots + key / num_shards) >> shift: out_regions gets the records bin after bin (capacity: the segment's positions),
// out_slots their virtual slots, bins_ctl (OSW_DIGITS + 1 words, the caller's) the bin sizes, counts_dev the records per destination
static int shard_lift_impl(wfb_engine_t *e, const wfb_functors_t *pre, const wfb_batch_t *batches_h, uint32_t nbatches, uint32_t num_shards,
                           void *out_regions, uint32_t region_capacity, uint32_t *counts_dev, void *stream,
                           uint32_t shard_slots, uint32_t shard_keys, uint32_t shift, uint32_t *out_slots, uint32_t *bins_ctl,
                           uint64_t *send_meta = nullptr, uint64_t watermark = 0);

int wfb_shard_lift(wfb_engine_t *e, const wfb_functors_t *pre, const wfb_batch_t *batches_h, uint32_t nbatches, uint32_t num_shards,
                   void *out_regions, uint32_t region_capacity, uint32_t *counts_dev, void *stream)
{
    return shard_lift_impl(e, pre, batches_h, nbatches, num_shards, out_regions, region_capacity, counts_dev, stream, 0, 0, 0, nullptr, nullptr);
}

static int shard_lift_impl(wfb_engine_t *e, const wfb_functors_t *pre, const wfb_batch_t *batches_h, uint32_t nbatches, uint32_t num_shards,
                           void *out_regions, uint32_t region_capacity, uint32_t *counts_dev, void *stream,
                           uint32_t shard_slots, uint32_t shard_keys, uint32_t shift, uint32_t *out_slots, uint32_t *bins_ctl,
                           uint64_t *send_meta, uint64_t watermark)
{
    // Map -> Filter -> lift in one streaming pass (no compaction chain: tile t owns positions [t*TILE, +TILE)), then one
    // stable partition pass on the destination (key % num_shards) that moves the lifted records into the shard regions.
    if (!e || !counts_dev || !out_regions || num_shards == 0 || num_shards > MAX_SHARDS || (nbatches && !batches_h)) return WFB_E_BADARG;
    const bool bucketed = shard_slots != 0;
    if (bucketed && (!out_slots || !bins_ctl || (shard_slots & (shard_slots - 1)) || static_cast<uint64_t>(num_shards) * shard_slots > 65536u ||
                     shard_keys > shard_slots || (shard_slots >> shift) == 0)) return WFB_E_BADARG;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int rc = e->ts.enter(s); if (rc) return rc;
    CK(cudaMemsetAsync(counts_dev, 0, sizeof(uint32_t) * (MAX_SHARDS + 1), s));
    std::vector<DevBatch> hb; hb.reserve(nbatches);
    uint64_t total = 0; uint32_t tiles = 0; uint64_t span_begin = ~0ull, span_end = 0;
    for (uint32_t i = 0; i < nbatches; i++) {
        if (batches_h[i].n == 0) continue;
        if (!batches_h[i].tuples) return WFB_E_BADARG;
        DevBatch b; std::memset(&b, 0, sizeof(b));
        b.tuples = static_cast<const unsigned char *>(batches_h[i].tuples); b.watermark = batches_h[i].watermark;
        b.n = batches_h[i].n; b.tile_begin = tiles;
        tiles += tiles_of(b.n); total += b.n;
        const uint64_t p0 = reinterpret_cast<uint64_t>(b.tuples);
        span_begin = std::min(span_begin, p0); span_end = std::max(span_end, p0 + static_cast<uint64_t>(b.n) * e->ops->tuple_bytes);
        hb.push_back(b);
    }
    if (total == 0) { if (bucketed) CK(cudaMemsetAsync(bins_ctl, 0, sizeof(uint32_t) * (OSW_DIGITS + 1), s)); return 0; }
    const uint64_t positions = static_cast<uint64_t>(tiles) * TILE;
    if (positions > 0x7fffffffull || (bucketed && positions > region_capacity)) return WFB_E_BADARG;
    nbatches = static_cast<uint32_t>(hb.size());
    const size_t RB = e->ops->result_bytes;
    if (positions > e->sh_cap) {
        CK(cudaStreamSynchronize(s));
        cudaFree(e->sh_lifted); cudaFree(e->sh_dest);
        e->sh_cap = std::max<uint64_t>(positions, 2 * e->sh_cap);
        CK(cudaMalloc(&e->sh_lifted, e->sh_cap * RB));
        CK(cudaMalloc(&e->sh_dest, e->sh_cap * sizeof(uint32_t)));
    }
    if (!e->sh_ctl) CK(cudaMalloc(&e->sh_ctl, sizeof(uint32_t) * RadixSorter::CTL_WORDS));
    rc = e->ts.ensure_tiles(tiles); if (rc) return rc;
    rc = e->ts.ensure_batches(nbatches); if (rc) return rc;
    { int rc_ = e->ts.stage.h2d(e->ts.d_batches, hb.data(), sizeof(DevBatch) * nbatches, s); if (rc_) return rc_; }
    uint32_t *ctl = bucketed ? bins_ctl : e->sh_ctl;
    rc = RadixSorter::prepare_wide(ctl, s); if (rc) return rc;
    TileArgs a; std::memset(&a, 0, sizeof(a));
    a.batches = e->ts.d_batches; a.nbatches = nbatches; a.num_tiles = tiles;
    a.lifted = e->sh_lifted; a.slots = e->sh_dest; a.nshards = num_shards; a.sparse = 1; a.l2_hints = 1;
    a.sort_ctl = ctl; a.sort_passes = 1; a.sort_shift = 0; a.sort_dbits = OSW_BITS;
    if (bucketed) { a.shard_slots = shard_slots; a.shard_keys = shard_keys; a.shard_err = counts_dev + MAX_SHARDS; a.sort_shift = shift; a.pack_rank = 1; }
    // the tile pass claims whole wide tiles and files the per-tile destination counts itself (no counting pass in the partition)
    static const bool h16_env = !(std::getenv("WFB_TILE_H16") && std::atoi(std::getenv("WFB_TILE_H16")) == 0);
    const bool h16 = h16_env || bucketed;
    uint32_t claims = tiles;
    if (h16) {
        rc = e->sorter.ensure_wide(static_cast<uint32_t>(positions), s, &a.wide_h16); if (rc) return rc;
        a.tiles_per_ticket = OSW_TILE_POS / TILE; a.sort_ctl = nullptr;
        claims = (tiles + a.tiles_per_ticket - 1) / a.tiles_per_ticket;
    }
    e->ts.next_launch(a);
    uint32_t grid = 0;
    rc = e->ops->tile_pass(MODE_INGEST, a, pre ? static_cast<const void *>(pre) : e->pp(), claims, s, &grid, span_begin, span_end); if (rc) return rc;
    e->ts.launched(claims, grid);
    e->launches++;
    const uint32_t *counts = nullptr;
    const uint64_t before = e->sorter.launches;
    if (bucketed) {
        rc = e->sorter.sort_wide<uint32_t>(e->sh_dest, out_slots, nullptr, nullptr, static_cast<uint32_t>(positions), static_cast<uint32_t>(positions), shift, s,
                                           ctl, &counts, e->sh_lifted, static_cast<unsigned char *>(out_regions), static_cast<uint32_t>(RB), true,
                                           0, 0, false, true, nullptr, true);
        if (rc) return rc;
        k_shard_bin_counts<<<1, 32 * MAX_SHARDS, 0, s>>>(counts, num_shards, shard_slots >> shift, counts_dev, send_meta, watermark);
    } else {
        rc = e->sorter.sort_wide<uint32_t>(e->sh_dest, nullptr, nullptr, nullptr, static_cast<uint32_t>(positions), static_cast<uint32_t>(positions), 0, s,
                                           e->sh_ctl, &counts, e->sh_lifted, static_cast<unsigned char *>(out_regions), static_cast<uint32_t>(RB), true,
                                           region_capacity, num_shards, false, h16);
        if (rc) return rc;
        k_shard_counts<<<1, 32, 0, s>>>(counts, num_shards, region_capacity, counts_dev);
    }
    CK(cudaGetLastError());
    e->launches += e->sorter.launches - before + 1;
    return 0;
}

// ---- Ffat_Windows_GPU ------------------------------------------------------------------------------------------
static uint64_t gcd_u64(uint64_t a, uint64_t b) { while (b) { uint64_t t = a % b; a = b; b = t; } return a; }

// ---- keyed-stateful Map_GPU / Filter_GPU ---------------------------------------------------------------------------------
} // extern "C" (the struct below is C++)
struct wfb_kstate {
    int prog = 0;
    const ProgramOps *ops = nullptr;
    FfatDev ff{};                 // key table only (dense or open addressing)
    unsigned char *states = nullptr;
    RadixSorter sorter;
    uint32_t bucket_shift = 0;
    DevBatch *d_batches = nullptr; uint32_t *d_boff = nullptr; uint32_t batch_cap = 0;
    uint32_t cap = 0;             // tuples per call
    uint32_t *slotsA = nullptr, *slotsB = nullptr, *posB = nullptr, *tile_cnt = nullptr, *rank_start = nullptr;
    unsigned char *keep = nullptr;
    uint64_t launches = 0;
    PinnedStage stage;
};
extern "C" {

int wfb_kstate_create(wfb_kstate_t **hh, int prog, uint32_t max_keys, uint32_t flags)
{
    if (!hh || max_keys == 0) return WFB_E_BADARG;
    const ProgramOps *o = program(prog);
    if (!o) return WFB_E_NOPROG;
    if (o->state_bytes == 0) return WFB_E_UNSUPPORTED; // the program has no state_t / stateful functors
    uint32_t bits = 0; while ((1ull << bits) < max_keys) bits++;
    if (bits > OSW_BITS + 6) return WFB_E_UNSUPPORTED;  // 1024 buckets of at most 64 keys
    int rc = device_ready(); if (rc) return rc;
    wfb_kstate *h = new (std::nothrow) wfb_kstate();
    if (!h) return WFB_E_BADARG;
    h->prog = prog; h->ops = o; h->bucket_shift = bits > OSW_BITS ? bits - OSW_BITS : 0;
    FfatDev &ff = h->ff;
    ff.max_keys = max_keys; ff.dense = (flags & WFB_FFAT_DENSE_KEYS) ? 1u : 0u;
    uint32_t cap = 1; while (cap < 2ull * max_keys) cap <<= 1;
    ff.ht_mask = cap - 1;
#define ALLOC(ptr, bytes) do { cudaError_t e_ = cudaMalloc(reinterpret_cast<void **>(&(ptr)), (bytes)); if (e_ != cudaSuccess) { wfb_kstate_destroy(h); return static_cast<int>(e_); } } while (0)
    if (!ff.dense) {
        ALLOC(ff.ht_keys, sizeof(uint64_t) * cap); ALLOC(ff.ht_slots, sizeof(uint32_t) * cap);
        CK(cudaMemset(ff.ht_keys, 0xff, sizeof(uint64_t) * cap)); CK(cudaMemset(ff.ht_slots, 0xff, sizeof(uint32_t) * cap));
    }
    ALLOC(ff.n_slots, sizeof(uint32_t) * 4); ff.err_flags = ff.n_slots + 1;
    CK(cudaMemset(ff.n_slots, 0, sizeof(uint32_t) * 4));
    ALLOC(ff.slot_key, sizeof(uint64_t) * max_keys);
    ALLOC(h->states, static_cast<size_t>(o->state_bytes) * max_keys);
    CK(cudaMemset(h->states, 0, static_cast<size_t>(o->state_bytes) * max_keys)); // state_t(): zero-initialised
#undef ALLOC
    *hh = h;
    return 0;
}

int wfb_kstate_destroy(wfb_kstate_t *h)
{
    if (!h) return 0;
    cudaDeviceSynchronize();
    cudaFree(h->ff.ht_keys); cudaFree(h->ff.ht_slots); cudaFree(h->ff.n_slots); cudaFree(h->ff.slot_key); cudaFree(h->states);
    cudaFree(h->d_batches); cudaFree(h->d_boff); cudaFree(h->slotsA); cudaFree(h->slotsB); cudaFree(h->posB); cudaFree(h->tile_cnt);
    cudaFree(h->rank_start); cudaFree(h->keep);
    h->sorter.destroy(); h->stage.destroy();
    delete h;
    cudaGetLastError();
    return 0;
}

static int kstate_run(wfb_kstate_t *h, const wfb_functors_t *f, const wfb_batch_t *in_h, const wfb_batch_t *out_h, uint32_t nbatches,
                      uint32_t *n_out_dev, bool filter, cudaStream_t s)
{
    if (!h || !f || (nbatches && !in_h) || (filter && (!out_h || !n_out_dev))) return WFB_E_BADARG;
    if (nbatches == 0) return 0;
    if (filter) CK(cudaMemsetAsync(n_out_dev, 0, sizeof(uint32_t) * nbatches, s));
    std::vector<DevBatch> hb(nbatches);
    std::vector<uint32_t> boff(nbatches + 1);
    uint64_t total = 0;
    for (uint32_t i = 0; i < nbatches; i++) {
        if (in_h[i].n && (!in_h[i].tuples || (filter && !out_h[i].tuples))) return WFB_E_BADARG;
        DevBatch b; std::memset(&b, 0, sizeof(b));
        b.tuples = static_cast<const unsigned char *>(in_h[i].tuples); b.ts = in_h[i].ts; b.n = in_h[i].n;
        if (filter) {
            b.out = static_cast<unsigned char *>(const_cast<void *>(out_h[i].tuples));
            b.ts_out = in_h[i].ts ? const_cast<uint64_t *>(out_h[i].ts) : nullptr;
            b.n_out = n_out_dev + i;
        }
        hb[i] = b; boff[i] = static_cast<uint32_t>(total); total += b.n;
    }
    boff[nbatches] = static_cast<uint32_t>(total);
    if (total == 0) return 0;
    if (total > 0x7fffffffull) return WFB_E_BADARG;
    const uint32_t n = static_cast<uint32_t>(total);
    if (nbatches > h->batch_cap) {
        CK(cudaStreamSynchronize(s));
        cudaFree(h->d_batches); cudaFree(h->d_boff); cudaFree(h->rank_start);
        h->batch_cap = std::max(nbatches, 2 * h->batch_cap);
        CK(cudaMalloc(&h->d_batches, sizeof(DevBatch) * h->batch_cap));
        CK(cudaMalloc(&h->d_boff, sizeof(uint32_t) * (static_cast<size_t>(h->batch_cap) + 1)));
        CK(cudaMalloc(&h->rank_start, sizeof(uint32_t) * (static_cast<size_t>(h->batch_cap) + 1)));
    }
    if (n > h->cap) {
        CK(cudaStreamSynchronize(s));
        cudaFree(h->slotsA); cudaFree(h->slotsB); cudaFree(h->posB); cudaFree(h->tile_cnt); cudaFree(h->keep);
        h->cap = std::max(n, 2 * h->cap);
        CK(cudaMalloc(&h->slotsA, sizeof(uint32_t) * h->cap)); CK(cudaMalloc(&h->slotsB, sizeof(uint32_t) * h->cap));
        CK(cudaMalloc(&h->posB, sizeof(uint32_t) * h->cap)); CK(cudaMalloc(&h->keep, h->cap));
        CK(cudaMalloc(&h->tile_cnt, sizeof(uint32_t) * ((h->cap + SEGT - 1) / SEGT + 1)));
    }
    { int rc_ = h->stage.h2d(h->d_batches, hb.data(), sizeof(DevBatch) * nbatches, s); if (rc_) return rc_; }
    { int rc_ = h->stage.h2d(h->d_boff, boff.data(), sizeof(uint32_t) * (nbatches + 1), s); if (rc_) return rc_; }
    // 1. slots; 2. one wide partition pass into 1024 buckets of consecutive slots; 3. per-bucket CTAs, one thread per key
    int rc = h->ops->ks_slots(h->d_batches, h->d_boff, nbatches, n, h->ff, h->slotsA, s, f); if (rc) return rc;
    const uint32_t *counts = nullptr;
    const uint64_t before = h->sorter.launches;
    rc = h->sorter.sort_wide<uint32_t>(h->slotsA, h->slotsB, h->posB, nullptr, n, n, h->bucket_shift, s, nullptr, &counts, nullptr, nullptr, 0, true);
    if (rc) return rc;
    rc = h->ops->ks_apply(filter ? 1 : 0, h->ff, h->d_batches, h->d_boff, nbatches, h->slotsB, h->posB, counts, h->bucket_shift, h->states,
                          h->keep, s, f);
    if (rc) return rc;
    h->launches += 2 + (h->sorter.launches - before);
    if (filter) { // stable per-batch compaction by the keep flags
        const uint32_t tiles = (n + SEGT - 1) / SEGT;
        k_flag_tile_counts<<<tiles, 256, 0, s>>>(h->keep, n, h->tile_cnt);
        k_scan_u32<<<1, 1024, 0, s>>>(h->tile_cnt, h->tile_cnt, tiles, nullptr);
        k_flag_batch_starts<<<(nbatches + 1 + 127) / 128, 128, 0, s>>>(h->keep, h->tile_cnt, h->d_boff, nbatches, n, h->rank_start, h->d_batches);
        k_flag_batch_counts<<<(nbatches + 127) / 128, 128, 0, s>>>(h->rank_start, nbatches, h->d_batches);
        CK(cudaGetLastError());
        rc = h->ops->flag_scatter(h->keep, h->tile_cnt, h->d_boff, nbatches, n, h->rank_start, h->d_batches, s); if (rc) return rc;
        h->launches += 5;
    }
    return 0;
}

int wfb_map_stateful(wfb_kstate_t *h, const wfb_functors_t *f, const wfb_batch_t *batches_h, uint32_t nbatches, void *stream)
{
    return kstate_run(h, f, batches_h, nullptr, nbatches, nullptr, false, static_cast<cudaStream_t>(stream));
}

int wfb_filter_stateful(wfb_kstate_t *h, const wfb_functors_t *f, const wfb_batch_t *in_h, const wfb_batch_t *out_h, uint32_t nbatches,
                        uint32_t *n_out_dev, void *stream)
{
    return kstate_run(h, f, in_h, out_h, nbatches, n_out_dev, true, static_cast<cudaStream_t>(stream));
}

static int ffat_process_cb_impl(wfb_ffat_t *h, const void *pre, const wfb_batch_t *batches_h, uint32_t nbatches, void *out_results, uint64_t *out_ts,
                                uint32_t out_capacity, uint32_t *n_out_dev, void *stream, const uint32_t *ext_slots);

// ---- time-based windows: front-end handle + count-based back end over the lifted program ------------------------------------
// id of the lifted variant of a program (LiftedOf<P>, registered on first use)
static int lifted_program_of(int prog)
{
    static std::vector<int> cache; // by program id
    static std::mutex cache_mutex;
    const ProgramOps *o = program(prog);
    if (!o || !o->lifted_ops) return -1;
    std::lock_guard<std::mutex> lock(cache_mutex);
    if (static_cast<size_t>(prog) < cache.size() && cache[prog] > 0) return cache[prog];
    const int id = wfb_program_register(o->lifted_ops(), sizeof(ProgramOps));
    if (id < 0) return id;
    if (cache.size() <= static_cast<size_t>(prog)) cache.resize(prog + 1, 0);
    cache[prog] = id;
    return id;
}

static int tb_create(wfb_ffat_t **hh, int prog, uint64_t win, uint64_t slide, uint32_t nb, uint32_t max_keys, uint64_t lateness, uint32_t flags)
{
    const ProgramOps *o = program(prog);
    if (!o) return WFB_E_NOPROG;
    const int lp = lifted_program_of(prog);
    if (lp < 0 || (flags & WFB_FFAT_PIPELINED)) return WFB_E_UNSUPPORTED;
    int rc = device_ready(); if (rc) return rc;
    const uint64_t pane_len = gcd_u64(win, slide);                   // wf/ffat_replica_gpu.hpp:639-642
    const uint64_t win_p = win / pane_len, slide_p = slide / pane_len;
    const uint64_t Bp = static_cast<uint64_t>(nb - 1) * slide_p + win_p, group = slide_p * nb;
    uint64_t capq = 1; while (capq < 2 * Bp + group + lateness / pane_len + 8) capq <<= 1;
    const size_t RB = o->result_bytes;
    if (Bp > (1ull << 24) || capq * max_keys * RB > (32ull << 30)) return WFB_E_BADARG;
    wfb_ffat *h = new (std::nothrow) wfb_ffat();
    if (!h) return WFB_E_BADARG;
    h->prog = prog; h->ops = o; h->win_type = 1; h->tb_lateness = lateness; h->pipelined = false;
    FfatDev &ff = h->ff;
    ff.max_keys = max_keys; ff.dense = (flags & WFB_FFAT_DENSE_KEYS) ? 1u : 0u; ff.nb = nb;
    uint32_t cap = 1; while (cap < 2ull * max_keys) cap <<= 1;
    ff.ht_mask = cap - 1;
    size_t total = 0;
#define ALLOC(ptr, bytes) do { cudaError_t e_ = cudaMalloc(reinterpret_cast<void **>(&(ptr)), (bytes)); if (e_ != cudaSuccess) { wfb_ffat_destroy(h); return static_cast<int>(e_); } total += (bytes); } while (0)
    if (!ff.dense) {
        ALLOC(ff.ht_keys, sizeof(uint64_t) * cap);
        ALLOC(ff.ht_slots, sizeof(uint32_t) * cap);
        CK(cudaMemset(ff.ht_keys, 0xff, sizeof(uint64_t) * cap));
        CK(cudaMemset(ff.ht_slots, 0xff, sizeof(uint32_t) * cap));
    }
    ALLOC(ff.n_slots, sizeof(uint32_t) * 4);
    h->own_n_slots = ff.n_slots;
    ff.err_flags = ff.n_slots + 1;
    CK(cudaMemset(ff.n_slots, 0, sizeof(uint32_t) * 4));
    ALLOC(ff.slot_key, sizeof(uint64_t) * max_keys);
    TbDev &tb = h->tb;
    tb.pane_len = pane_len; tb.Bp = Bp; tb.group = group; tb.capq = static_cast<uint32_t>(capq);
    ALLOC(tb.first, sizeof(uint64_t) * max_keys); CK(cudaMemset(tb.first, 0, sizeof(uint64_t) * max_keys));
    ALLOC(tb.num, sizeof(uint32_t) * max_keys); CK(cudaMemset(tb.num, 0, sizeof(uint32_t) * max_keys));
    ALLOC(tb.num_new, sizeof(uint32_t) * max_keys);
    ALLOC(tb.trig, sizeof(uint64_t) * max_keys);
    { std::vector<uint64_t> t0(max_keys, Bp - 1); CK(cudaMemcpy(tb.trig, t0.data(), sizeof(uint64_t) * max_keys, cudaMemcpyHostToDevice)); } // :463
    ALLOC(tb.done, sizeof(uint32_t) * max_keys); CK(cudaMemset(tb.done, 0, sizeof(uint32_t) * max_keys));
    ALLOC(tb.ring, RB * capq * max_keys);
    ALLOC(tb.present, sizeof(uint32_t) * max_keys);
    ALLOC(tb.cnt, sizeof(uint32_t) * (static_cast<size_t>(max_keys) + 1));
    ALLOC(h->tb_misc, sizeof(uint32_t) * 8); CK(cudaMemset(h->tb_misc, 0, sizeof(uint32_t) * 8));
    tb.n_present = h->tb_misc + 2; tb.ignored = h->tb_misc + 4; tb.need = h->tb_misc + 5; tb.err = ff.err_flags;
#undef ALLOC
    rc = h->ts.init(); if (rc) { wfb_ffat_destroy(h); return rc; }
    rc = wfb_ffat_create(&h->cb, lp, win_p, slide_p, nb, max_keys, 0, 0, flags & WFB_FFAT_DENSE_KEYS);
    if (rc) { wfb_ffat_destroy(h); return rc; }
    // the front end hands the popped panes to the back end in place, with the slot of every record: that needs the bucket path with the
    // in-place ingest (at most 65536 keys; not with the WFB_SPARSE=0 / WFB_INPLACE=0 / WFB_UPDATE=lanes / WFB_BUCKET_MOVE=1 knobs). Refuse here,
    // before any pane has been consumed, rather than at the first firing batch
    if (!h->cb->buckets || !h->cb->sparse_ingest || !h->cb->inplace_ok || h->cb->bucket_move) { wfb_ffat_destroy(h); return WFB_E_UNSUPPORTED; }
    // the back end never looks keys up (the front end hands it the slot of every record): it only needs slot -> key for the results
    if (!ff.dense) { cudaFree(h->cb->ff.slot_key); h->cb->ff.slot_key = ff.slot_key; h->cb->ff.n_slots = ff.n_slots; h->cb->shares_slot_key = true; }
    h->state_bytes = total + h->cb->state_bytes;
    *hh = h;
    return 0;
}

static int tb_ensure(wfb_ffat *h, uint32_t n, cudaStream_t s)
{
    if (n <= h->tb_cap) return 0;
    CK(cudaStreamSynchronize(s));
    cudaFree(h->tb_kA); cudaFree(h->tb_kB); cudaFree(h->tb_iA); cudaFree(h->tb_iB); cudaFree(h->tb_lifted); cudaFree(h->tb_part);
    cudaFree(h->tb_head); cudaFree(h->tb_seg);
    h->tb_cap = std::max(n, 2 * h->tb_cap);
    const size_t RB = h->ops->result_bytes, c = h->tb_cap;
    CK(cudaMalloc(&h->tb_kA, sizeof(uint64_t) * c)); CK(cudaMalloc(&h->tb_kB, sizeof(uint64_t) * c));
    CK(cudaMalloc(&h->tb_iA, sizeof(uint32_t) * c)); CK(cudaMalloc(&h->tb_iB, sizeof(uint32_t) * c));
    CK(cudaMalloc(&h->tb_lifted, RB * c)); CK(cudaMalloc(&h->tb_part, RB * c));
    CK(cudaMalloc(&h->tb_head, sizeof(uint32_t) * ((c + SEGT - 1) / SEGT + 1))); CK(cudaMalloc(&h->tb_seg, sizeof(uint32_t) * (c + 1)));
    return 0;
}

int wfb_ffat_process_tb(wfb_ffat_t *h, const wfb_functors_t *pre, const wfb_batch_t *batches_h, uint32_t nbatches,
                        void *out_results, uint64_t *out_ts, uint32_t out_capacity, uint32_t *n_out_dev, void *stream)
{
    if (!h || h->win_type != 1 || !n_out_dev || (nbatches && !batches_h) || (out_capacity && !out_results)) return WFB_E_BADARG;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t RB = h->ops->result_bytes;
    const void *prm = pre ? static_cast<const void *>(pre) : h->pp();
    uint32_t produced = 0;
    for (uint32_t bi = 0; bi < nbatches; bi++) {
        const wfb_batch_t &b = batches_h[bi];
        if (b.n == 0) continue;
        if (!b.tuples || !b.ts) return WFB_E_BADARG;            // time-based windows need the timestamps
        int rc = tb_ensure(h, b.n, s); if (rc) return rc;
        const uint64_t wm = b.watermark;
        const uint64_t F = wm >= h->tb_lateness ? (wm - h->tb_lateness) / h->tb.pane_len : 0; // first_pane_not_complete :875-881
        uint32_t *misc = h->tb_misc;
        CK(cudaMemsetAsync(misc, 0, sizeof(uint32_t) * 4, s));
        CK(cudaMemsetAsync(misc + 5, 0, sizeof(uint32_t), s));
        // 1. lift + composite (slot, pane) keys; 2. stable sort; 3. (key, pane) segments; 4. partials; 5. merge into the rings
        rc = h->ops->tb_lift(static_cast<const unsigned char *>(b.tuples), b.ts, b.n, h->ff, h->tb, F, h->tb_lifted, h->tb_kA, s, prm); if (rc) return rc;
        uint32_t tb_need_bits = 2;
        { // the rings must hold every pane from a key's first pending one to its newest (PendingPanes_Queue::push_panes :367-372)
            uint32_t need = 0;
            CK(cudaMemcpyAsync(&need, misc + 5, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
            CK(cudaStreamSynchronize(s));
            tb_need_bits = std::max(2u, need);
            if (need > h->tb.capq) {
                uint64_t ncap = h->tb.capq; while (ncap < need) ncap <<= 1;
                if (ncap * h->ff.max_keys * RB > (32ull << 30)) return WFB_E_CAPACITY;
                unsigned char *nr = nullptr;
                CK(cudaMalloc(&nr, RB * ncap * h->ff.max_keys));
                k_tb_ring_resize<<<grid_for(h->ff.max_keys, 128), 128, 0, s>>>(h->tb, h->tb.ring, h->tb.capq, nr, static_cast<uint32_t>(ncap), h->ff.max_keys,
                                                                                static_cast<uint32_t>(RB));
                CK(cudaStreamSynchronize(s));
                cudaFree(h->tb.ring);
                h->tb.ring = nr; h->tb.capq = static_cast<uint32_t>(ncap);
                h->launches++;
            }
        }
        // sort keys (slot << kbits) | relative pane, with just the bits this batch needs; filtered tuples sort behind every slot
        uint32_t sbits = 0; while ((1ull << sbits) < static_cast<uint64_t>(h->ff.max_keys) + 1) sbits++;
        uint32_t kbits = 1; while ((1ull << kbits) < tb_need_bits) kbits++;
        h->tb.kbits = kbits;
        k_tb_pack<<<grid_for(b.n, 256), 256, 0, s>>>(h->tb_kA, b.n, kbits, h->ff.max_keys);
        const uint64_t *skeys; const uint32_t *sidx;
        const uint64_t before = h->sorter.launches;
        rc = h->sorter.sort<uint64_t>(h->tb_kA, h->tb_kB, h->tb_iA, h->tb_iB, nullptr, b.n, b.n, (kbits + sbits + 7) / 8, s, &skeys, &sidx); if (rc) return rc;
        const uint32_t tiles = (b.n + SEGT - 1) / SEGT;
        k_head_tile_counts<<<tiles, 256, 0, s>>>(skeys, b.n, h->tb_head);
        k_scan_u32<<<1, 1024, 0, s>>>(h->tb_head, h->tb_head, tiles, nullptr);
        k_seg_finish_batches<<<tiles, 256, 0, s>>>(skeys, b.n, 64u, h->tb_head, h->tb_seg, misc + 1, misc + 0);
        CK(cudaGetLastError());
        rc = h->ops->tb_reduce(h->tb_lifted, skeys, sidx, h->tb_seg, misc + 0, h->tb_part, b.n, kbits, h->ff.max_keys, s, prm); if (rc) return rc;
        rc = h->ops->tb_merge(skeys, h->tb_seg, misc + 0, h->tb_part, h->ff, h->tb, b.n, s, prm); if (rc) return rc;
        // 6. panes to pop per present key, offsets, total
        const uint32_t maxp = std::min<uint32_t>(b.n, h->ff.max_keys);
        k_tb_pop_count<<<grid_for(maxp, 128), 128, 0, s>>>(h->tb, F);
        k_tb_scan_present<<<1, 1024, 0, s>>>(h->tb.cnt, h->tb.n_present, misc + 3);
        CK(cudaGetLastError());
        uint32_t total = 0;
        CK(cudaMemcpyAsync(&total, misc + 3, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));                           // (the reference synchronises here as well, :962)
        if (total > h->tb_pop_cap) {
            cudaFree(h->tb_popped); cudaFree(h->tb_popped_slots);
            h->tb_pop_cap = std::max(total, 2 * h->tb_pop_cap);
            CK(cudaMalloc(&h->tb_popped, RB * h->tb_pop_cap));
            CK(cudaMalloc(&h->tb_popped_slots, sizeof(uint32_t) * ((static_cast<size_t>(h->tb_pop_cap) + TILE - 1) / TILE * TILE)));
        }
        rc = h->ops->tb_pop_write(h->ff, h->tb, F, h->tb.cnt, h->tb_popped, h->tb_popped_slots, h->tb_pop_cap, maxp, s, prm); if (rc) return rc;
        h->launches += 10 + (h->sorter.launches - before);
        // 7. the count-based back end consumes the popped panes as one batch with this batch's watermark
        if (total) {
            wfb_batch_t pb; std::memset(&pb, 0, sizeof(pb));
            pb.tuples = h->tb_popped; pb.ts = nullptr; pb.watermark = wm; pb.n = total;
            const uint64_t lb = h->cb->launches;
            rc = ffat_process_cb_impl(h->cb, prm, &pb, 1, static_cast<unsigned char *>(out_results) + static_cast<size_t>(produced) * RB,
                                      out_ts ? out_ts + produced : nullptr, out_capacity - produced, n_out_dev, s, h->tb_popped_slots);
            if (rc) return rc;
            h->launches += h->cb->launches - lb;
            uint32_t got = 0;
            CK(cudaMemcpyAsync(&got, n_out_dev, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
            CK(cudaStreamSynchronize(s));
            produced += std::min(got, out_capacity - produced);
        }
    }
    CK(cudaMemcpyAsync(n_out_dev, &produced, sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    CK(cudaStreamSynchronize(s)); // `produced` lives on this stack frame
    return 0;
}

int wfb_ffat_create(wfb_ffat_t **hh, int prog, uint64_t win, uint64_t slide, uint32_t wins_per_batch,
                    uint32_t max_keys, int win_type, uint64_t lateness, uint32_t flags)
{
    if (!hh || win == 0 || slide == 0 || wins_per_batch == 0 || max_keys == 0) return WFB_E_BADARG;
    if (win_type == 1) return tb_create(hh, prog, win, slide, wins_per_batch, max_keys, lateness, flags);
    if (win_type != 0) return WFB_E_UNSUPPORTED;
    const ProgramOps *o = program(prog);
    if (!o) return WFB_E_NOPROG;
    int rc = device_ready(); if (rc) return rc;
    wfb_ffat *h = new (std::nothrow) wfb_ffat();
    if (!h) return WFB_E_BADARG;
    h->prog = prog; h->ops = o; h->win_type = win_type;
    { const char *e = std::getenv("WFB_SORT_PAYLOAD"); h->move_payload = e && std::atoi(e) != 0; }
    rc = h->ts.init(); if (rc) { delete h; return rc; }
    FfatDev &ff = h->ff;
    ff.win = win; ff.slide = slide; ff.nb = wins_per_batch;
    ff.B = static_cast<uint64_t>(wins_per_batch - 1) * slide + win;
    const uint64_t pane = gcd_u64(win, slide);
    const uint64_t bp = ff.B / pane;
    if (pane > 0xffffffffull || bp > (1ull << 30)) { delete h; return WFB_E_BADARG; }
    ff.pane = static_cast<uint32_t>(pane); ff.wp = static_cast<uint32_t>(win / pane); ff.sp = static_cast<uint32_t>(slide / pane);
    // ring of n leaves (a power of two) for the bp panes a group reads + spare leaves: a key that completes up to `spare` further panes in
    // the call that fires a group leaves the group's leaves alone, so the group can wait for the deferred pass (one warp per group, levels
    // built on chip) instead of being evaluated inside the update kernel. WFB_RING_SPARE: minimum spare leaves (default min(bp, 32)).
    static const int spare_env = std::getenv("WFB_RING_SPARE") ? std::atoi(std::getenv("WFB_RING_SPARE")) : -1;
    const uint64_t spare_min = spare_env >= 0 ? static_cast<uint64_t>(spare_env) : std::min<uint64_t>(bp, 32);
    uint32_t n = 1, lg = 0; while (n < bp + spare_min) { n <<= 1; lg++; }
    ff.n_leaves = n; ff.log_leaves = lg;
    ff.defer_items = (static_cast<uint64_t>(n) - bp + 1) * pane;
    ff.max_keys = max_keys; ff.dense = (flags & WFB_FFAT_DENSE_KEYS) ? 1u : 0u;
    uint32_t cap = 1; while (cap < 2ull * max_keys) cap <<= 1;
    ff.ht_mask = cap - 1;
    const size_t RB = o->result_bytes;
    const size_t tree_bytes = static_cast<size_t>(max_keys) * (2ull * n - 1) * RB;
    size_t total = 0;
#define ALLOC(ptr, bytes) do { cudaError_t e_ = cudaMalloc(reinterpret_cast<void **>(&(ptr)), (bytes)); if (e_ != cudaSuccess) { wfb_ffat_destroy(h); return static_cast<int>(e_); } total += (bytes); } while (0)
    if (!ff.dense) {
        ALLOC(ff.ht_keys, sizeof(uint64_t) * cap);
        ALLOC(ff.ht_slots, sizeof(uint32_t) * cap);
        CK(cudaMemset(ff.ht_keys, 0xff, sizeof(uint64_t) * cap));
        CK(cudaMemset(ff.ht_slots, 0xff, sizeof(uint32_t) * cap));
    }
    ALLOC(ff.n_slots, sizeof(uint32_t) * 4);
    h->own_n_slots = ff.n_slots;
    ff.err_flags = ff.n_slots + 1;
    ff.results_total = reinterpret_cast<unsigned long long *>(ff.n_slots + 2);
    CK(cudaMemset(ff.n_slots, 0, sizeof(uint32_t) * 4));
    ALLOC(ff.slot_key, sizeof(uint64_t) * max_keys);
    ALLOC(ff.cnt, sizeof(uint64_t) * max_keys);
    CK(cudaMemset(ff.cnt, 0, sizeof(uint64_t) * max_keys));
    ALLOC(ff.acc, RB * max_keys);
    CK(cudaMemset(ff.acc, 0, RB * max_keys));
    ALLOC(ff.tree, tree_bytes);
    CK(cudaMemset(ff.tree, 0, tree_bytes));
    ALLOC(ff.seg_off, sizeof(uint32_t) * (static_cast<size_t>(max_keys) + 1));
    CK(cudaMemset(ff.seg_off, 0xff, sizeof(uint32_t) * (static_cast<size_t>(max_keys) + 1)));
    ALLOC(ff.heavy, sizeof(uint32_t) * max_keys);
    { const char *e = std::getenv("WFB_LIGHT_MAX"); ff.light_max = e ? static_cast<uint32_t>(std::atoi(e)) : 256u; }
    h->pipelined = (flags & WFB_FFAT_PIPELINED) != 0;
    for (int p = 0; p < (h->pipelined ? 2 : 1); p++) {
        SegScratch &g = h->seg[p];
        ALLOC(g.seg_cnt, sizeof(uint32_t) * max_keys);
        CK(cudaMemset(g.seg_cnt, 0, sizeof(uint32_t) * max_keys));
        ALLOC(g.n_total, sizeof(uint32_t) * 4);
        CK(cudaMemset(g.n_total, 0, sizeof(uint32_t) * 4));
        g.n_trig = g.n_total + 1; g.res_n = nullptr; g.n_heavy = g.n_total + 3;
        ALLOC(g.sort_ctl, sizeof(uint32_t) * RadixSorter::CTL_WORDS);
        CK(cudaEventCreateWithFlags(&g.ev_ingest, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&g.ev_done, cudaEventDisableTiming));
    }
    if (h->pipelined) {
        CK(cudaStreamCreateWithFlags(&h->s2, cudaStreamNonBlocking));
        const char *e = std::getenv("WFB_INGEST_CTAS_PER_SM");
        h->ingest_ctas_per_sm = e ? static_cast<uint32_t>(std::atoi(e)) : 2u;
    }
#undef ALLOC
    uint32_t bits = 0; while ((1ull << bits) < max_keys) bits++;
    h->sort_passes = std::max(1u, (bits + 7) / 8);
    h->bucket_shift = bits > OSW_BITS ? bits - OSW_BITS : 0;
    { const char *e = std::getenv("WFB_BUCKET_MOVE"); h->bucket_move = e && std::atoi(e) != 0; }
    { const char *e = std::getenv("WFB_L2_HINTS"); h->l2_hints = !(e && std::atoi(e) == 0); }
    { const char *e = std::getenv("WFB_SPARSE"); h->sparse_ingest = !(e && std::atoi(e) == 0); }
    { const char *e = std::getenv("WFB_INPLACE"); h->inplace_ok = !(e && std::atoi(e) == 0); }
    { const char *e = std::getenv("WFB_INPLACE_KERNEL"); h->inplace_kernel = !(e && std::atoi(e) == 0); }
    { const char *e = std::getenv("WFB_FUSE_TILE_HIST"); h->fuse_tile_hist = e && std::atoi(e) != 0; }
    { // WFB_L2_PERSIST=<MB>: L2 set-aside for evict-last lines (the lifted records between the ingest pass and the update)
        const char *e = std::getenv("WFB_L2_PERSIST");
        if (e && std::atoi(e) > 0) {
            int dev = 0, maxp = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&maxp, cudaDevAttrMaxPersistingL2CacheSize, dev);
            const size_t want = std::min<size_t>(static_cast<size_t>(std::atoi(e)) << 20, static_cast<size_t>(maxp));
            cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want);
            if (std::getenv("WFB_VERBOSE")) std::fprintf(stderr, "[wfb] persisting L2: max %d MB, set %zu MB\n", maxp >> 20, want >> 20);
        }
    }
    { // bucket path: every bucket holds at most BK_KEYS keys and the pane length fits 32 bits
        const char *e = std::getenv("WFB_UPDATE");
        h->buckets = !(e && std::strcmp(e, "lanes") == 0) && (1u << h->bucket_shift) <= BK_KEYS && ff.pane < (1ull << 32);
        // streaming update: one path update per completed pane inside the item loop, so panes of a few items at least
        // lazy FlatFAT levels (FfatDev::lazy): bucket path, the on-chip tree of one group must fit 32 KB, and building the n - 1 internal
        // nodes once per fired group must be cheaper than a root path (log n nodes) per completed pane: a group fires every sp * Nb panes.
        // (Nb = 1 with slide = pane fires on every pane: eager levels there. WFB_LAZY_TREE=0 / 1 forces either.)
        { const char *lz = std::getenv("WFB_LAZY_TREE");
          const bool fits = h->buckets && static_cast<size_t>(2) * ff.n_leaves * RB <= (32u << 10);
          const bool pays = static_cast<uint64_t>(ff.n_leaves) <= 2ull * std::max(1u, ff.log_leaves) * ff.sp * ff.nb;
          ff.lazy = (fits && (lz ? std::atoi(lz) != 0 : pays)) ? 1u : 0u; }
        h->stream_update = h->buckets && !h->bucket_move && e && std::strcmp(e, "stream") == 0; // (measured: 173 us against 153 us for the bucket kernel at the bench configuration)
        const char *t = std::getenv("WFB_TILE_H16");
        h->tile_h16 = h->buckets && !(t && std::atoi(t) == 0);
        const char *rs = std::getenv("WFB_RANK_SCATTER");
        h->rank_scatter = !(rs && std::atoi(rs) == 0);
    }
    h->state_bytes = total;
    *hh = h;
    return 0;
}

int wfb_ffat_destroy(wfb_ffat_t *h)
{
    if (!h) return 0;
    cudaDeviceSynchronize();
    FfatDev &ff = h->ff;
    cudaFree(ff.ht_keys); cudaFree(ff.ht_slots); cudaFree(h->own_n_slots); if (!h->shares_slot_key) cudaFree(ff.slot_key); cudaFree(ff.cnt);
    cudaFree(ff.acc); cudaFree(ff.tree); cudaFree(ff.seg_off); cudaFree(ff.heavy); cudaFree(h->mg_scratch);
    for (int p = 0; p < 2; p++) h->seg[p].destroy();
    h->sorter.destroy();
    if (h->cb) wfb_ffat_destroy(h->cb);
    cudaFree(h->tb.first); cudaFree(h->tb.num); cudaFree(h->tb.num_new); cudaFree(h->tb.trig); cudaFree(h->tb.done); cudaFree(h->tb.ring);
    cudaFree(h->tb.present); cudaFree(h->tb.cnt); cudaFree(h->tb_misc);
    cudaFree(h->tb_kA); cudaFree(h->tb_kB); cudaFree(h->tb_iA); cudaFree(h->tb_iB); cudaFree(h->tb_lifted); cudaFree(h->tb_part);
    cudaFree(h->tb_popped); cudaFree(h->tb_popped_slots); cudaFree(h->tb_head); cudaFree(h->tb_seg);
    if (h->s2) cudaStreamDestroy(h->s2);
    for (auto &e : h->tev) cudaEventDestroy(e);
    h->ts.destroy();
    delete h;
    cudaGetLastError();
    return 0;
}

uint64_t wfb_ffat_launches(const wfb_ffat_t *h) { return h ? h->launches : 0; }

int wfb_ffat_set_params(wfb_ffat_t *h, const void *params, size_t bytes)
{
    if (!h || !params || bytes != h->ops->params_bytes) return WFB_E_BADARG;
    h->params.assign(static_cast<const unsigned char *>(params), static_cast<const unsigned char *>(params) + bytes);
    if (h->cb) h->cb->params = h->params; // the lifted variant shares the program's params_t (its comb / make_result)
    return 0;
}
uint64_t wfb_ffat_state_bytes(const wfb_ffat_t *h) { return h ? h->state_bytes : 0; }
int wfb_ffat_set_key_shard(wfb_ffat_t *h, uint32_t num_shards, uint32_t shard)
{
    if (!h || num_shards == 0 || shard >= num_shards || !h->ff.dense || h->call_no != 0) return WFB_E_BADARG;
    if (h->win_type == 1) { // time-based: the front end maps keys to slots, the back end turns slots back into keys for the results
        if (h->cb == nullptr || h->cb->call_no != 0) return WFB_E_BADARG;
        h->cb->ff.key_div = num_shards; h->cb->ff.key_rem = shard;
    }
    h->ff.key_div = num_shards; h->ff.key_rem = shard;
    return 0;
}

static double host_now_us() { timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec * 1e6 + t.tv_nsec * 1e-3; }
static double g_sec[8]; static double g_sec_t = 0; static bool g_sec_on = false;
#define SEC(i) do { if (g_sec_on) { const double n_ = host_now_us(); g_sec[i] += n_ - g_sec_t; g_sec_t = n_; } } while (0)
static int ffat_ensure_segment(wfb_ffat *h, SegScratch &g, uint32_t total, uint32_t nbatches, cudaStream_t s)
{
    if (total > g.cap) {
        CK(cudaStreamSynchronize(s));
        if (h->s2) CK(cudaStreamSynchronize(h->s2));
        cudaFree(g.lifted); cudaFree(g.lifted_sorted); cudaFree(g.slotsA); cudaFree(g.slotsB); cudaFree(g.posA); cudaFree(g.posB);
        cudaFree(g.trig); cudaFree(g.res); cudaFree(g.res_ts);
        g.cap = std::max(total, 2 * g.cap);
        const size_t RB = h->ops->result_bytes;
        CK(cudaMalloc(&g.lifted, static_cast<size_t>(g.cap) * RB));
        if (h->move_payload || h->bucket_move) CK(cudaMalloc(&g.lifted_sorted, static_cast<size_t>(g.cap) * RB));
        CK(cudaMalloc(&g.slotsA, sizeof(uint32_t) * g.cap)); CK(cudaMalloc(&g.slotsB, sizeof(uint32_t) * g.cap));
        CK(cudaMalloc(&g.posA, sizeof(uint32_t) * g.cap)); CK(cudaMalloc(&g.posB, sizeof(uint32_t) * g.cap));
        const uint64_t per_group = std::max<uint64_t>(1, h->ff.slide * h->ff.nb);
        g.trig_cap = static_cast<uint32_t>(std::min<uint64_t>(g.cap / per_group + h->ff.max_keys + 1, 0x7fffffffull));
        CK(cudaMalloc(&g.trig, sizeof(Trigger) * g.trig_cap));
        if (h->pipelined) { // every group that can fire in one segment: trig_cap groups of Nb results
            g.res_cap = static_cast<uint32_t>(std::min<uint64_t>(static_cast<uint64_t>(g.trig_cap) * h->ff.nb, 0x7fffffffull));
            CK(cudaMalloc(&g.res, static_cast<size_t>(g.res_cap) * RB));
            CK(cudaMalloc(&g.res_ts, sizeof(uint64_t) * g.res_cap));
            g.res_n = g.n_total + 2;
        }
    }
    if (nbatches + 1 > g.batch_off_cap) {
        CK(cudaStreamSynchronize(s));
        if (h->s2) CK(cudaStreamSynchronize(h->s2));
        cudaFree(g.batch_off); cudaFree(g.d_batches);
        g.batch_off_cap = std::max(nbatches + 1, 2 * g.batch_off_cap);
        CK(cudaMalloc(&g.batch_off, sizeof(uint32_t) * g.batch_off_cap));
        CK(cudaMalloc(&g.d_batches, sizeof(DevBatch) * g.batch_off_cap));
    }
    return 0;
}

// sort + update + deferred window queries of the segment held in `g`, results to (out, out_ts, n_out)
static int ffat_window_phase(wfb_ffat *h, SegScratch &g, const FfatDev &ff, unsigned char *out, uint64_t *out_ts, uint32_t out_cap,
                             uint32_t *n_out, cudaStream_t s)
{
    const uint32_t *sorted_slots, *sorted_pos;
    const uint64_t before = h->sorter.launches;
    int rc;
    if (h->buckets) {
        // ONE wide radix pass on the top 10 slot bits: 1024 buckets of consecutive keys, arrival order inside a bucket ...
        const uint32_t *counts = nullptr;
        rc = h->sorter.sort_wide<uint32_t>(g.slotsA, g.slotsB, g.posB, g.sparse ? nullptr : g.n_total, g.total, g.total, h->bucket_shift, s,
                                           g.hist_ready ? g.sort_ctl : nullptr, &counts, h->bucket_move ? g.lifted : nullptr,
                                           h->bucket_move ? g.lifted_sorted : nullptr, static_cast<uint32_t>(h->ops->result_bytes), g.sparse, 0, 0, g.h32_ready, g.h16_ready, h->pipelined ? g.h16 : nullptr, g.ranked);
        if (rc) return rc;
        h->launches += h->sorter.launches - before;
        h->mark(2, s);
        SEC(3);
        // ... then one CTA per bucket finishes the job (local split by key, per-key ordered fold, FlatFAT update)
        if (h->stream_update && h->ops->ffat_stream)
            rc = h->ops->ffat_stream(ff, g.lifted_src, g.slotsB, g.posB, counts, h->bucket_shift, g.batch_off, g.d_batches, g.nbatches, out, out_ts, out_cap, n_out, s, h->pp());
        else
            rc = h->ops->ffat_buckets(ff, h->bucket_move ? g.lifted_sorted : g.lifted_src, g.slotsB, g.posB, counts, h->bucket_shift, h->bucket_move ? 1u : 0u, g.batch_off, g.d_batches, g.nbatches, out, out_ts,
                                      out_cap, n_out, s, h->pp());
        if (rc) return rc;
        h->launches += 1;
    } else {
        // stable sort of (slot, arrival position) by slot: onesweep radix, 8 bits per pass
        rc = h->sorter.sort<uint32_t>(g.slotsA, g.slotsB, g.posA, g.posB, g.n_total, 0, g.total, h->sort_passes, s, &sorted_slots,
                                      &sorted_pos, h->move_payload ? g.lifted : nullptr, h->move_payload ? g.lifted_sorted : nullptr,
                                      h->ops->result_bytes, g.hist_ready ? g.sort_ctl : nullptr, ff.seg_off, ff.max_keys);
        if (rc) return rc;
        h->launches += h->sorter.launches - before;
        h->mark(2, s);
        // one thread per key (one warp per heavy key): pane fold, FlatFAT update
        uint32_t ugrid = std::max(1u, std::min((ff.max_keys + 7) / 8, static_cast<uint32_t>(g_num_sms) * 8u));
        rc = h->ops->ffat_update(ff, h->move_payload ? g.lifted_sorted : g.lifted, sorted_pos, g.batch_off, g.d_batches, g.nbatches,
                                 out, out_ts, out_cap, n_out, ugrid, s, h->move_payload ? 0u : 1u, h->pp(),
                                 h->ff.light_max ? std::max(1u, std::min((ff.max_keys + 127u) / 128u, static_cast<uint32_t>(g_num_sms) * 16u)) : 0u);
        if (rc) return rc;
        h->launches += h->ff.light_max ? 2 : 1;
    }
    SEC(4);
    // deferred window groups: one thread per window
    rc = h->ops->ffat_windows(ff, g.batch_off, g.d_batches, g.nbatches, out, out_ts, out_cap, static_cast<uint32_t>(g_num_sms) * 4u, s, h->pp(), n_out);
    if (rc) return rc;
    h->launches += 1;
    return 0;
}

// pipelined mode: hand the finished results of the segment in `g` to the caller (stream-ordered on s)
static int ffat_deliver(wfb_ffat *h, SegScratch &g, unsigned char *out, uint64_t *out_ts, uint32_t out_cap, uint32_t *n_out, cudaStream_t s)
{
    if (!g.pending) { CK(cudaMemsetAsync(n_out, 0, sizeof(uint32_t), s)); return 0; }
    CK(cudaStreamWaitEvent(s, g.ev_done, 0));
    k_copy_results<<<static_cast<uint32_t>(g_num_sms) * 2u, 256, 0, s>>>(g.res, g.res_ts, g.res_n, h->ops->result_bytes, out, out_ts,
                                                                          out_cap, n_out, h->ff.err_flags);
    CK(cudaGetLastError());
    h->launches++;
    g.pending = false;
    return 0;
}

int wfb_ffat_process_cb(wfb_ffat_t *h, const wfb_functors_t *pre, const wfb_batch_t *batches_h, uint32_t nbatches,
                        void *out_results, uint64_t *out_ts, uint32_t out_capacity, uint32_t *n_out_dev, void *stream)
{
    return ffat_process_cb_impl(h, pre, batches_h, nbatches, out_results, out_ts, out_capacity, n_out_dev, stream, nullptr);
}

// ext_slots != nullptr: the key slot of the record at every position is given (one batch, read in place; used by the
// time-based front end, whose programs' lifted variants have no key extractor)
static int ffat_process_cb_impl2(wfb_ffat_t *h, const void *pre, const wfb_batch_t *batches_h, uint32_t nbatches,
                                 void *out_results, uint64_t *out_ts, uint32_t out_capacity, uint32_t *n_out_dev, void *stream,
                                 const uint32_t *ext_slots);
static int ffat_process_cb_impl(wfb_ffat_t *h, const void *pre, const wfb_batch_t *batches_h, uint32_t nbatches,
                                void *out_results, uint64_t *out_ts, uint32_t out_capacity, uint32_t *n_out_dev, void *stream,
                                const uint32_t *ext_slots)
{
    static const bool prof = std::getenv("WFB_HOST_PROFILE") != nullptr; // host time spent issuing a call (tuning aid)
    if (!prof) return ffat_process_cb_impl2(h, pre, batches_h, nbatches, out_results, out_ts, out_capacity, n_out_dev, stream, ext_slots);
    static double acc = 0; static uint64_t calls = 0;
    const double t0 = host_now_us();
    g_sec_on = true; g_sec_t = t0;
    const int rc = ffat_process_cb_impl2(h, pre, batches_h, nbatches, out_results, out_ts, out_capacity, n_out_dev, stream, ext_slots);
    acc += host_now_us() - t0;
    if (++calls % 64 == 0) {
        std::fprintf(stderr, "[wfb] process_cb host issue: %.1f us/call over the last 64 calls; sections:", acc / 64); acc = 0;
        for (int i = 0; i < 8; i++) { std::fprintf(stderr, " %.1f", g_sec[i] / 64); g_sec[i] = 0; }
        std::fprintf(stderr, "\n");
    }
    return rc;
}
static int ffat_process_cb_impl2(wfb_ffat_t *h, const void *pre, const wfb_batch_t *batches_h, uint32_t nbatches,
                                 void *out_results, uint64_t *out_ts, uint32_t out_capacity, uint32_t *n_out_dev, void *stream,
                                 const uint32_t *ext_slots)
{
    if (!h || !n_out_dev || (nbatches && !batches_h) || (out_capacity && !out_results)) return WFB_E_BADARG;
    if (h->win_type != 0) return WFB_E_BADARG; // time-based handles: wfb_ffat_process_tb
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int rc = h->ts.enter(s); if (rc) return rc;
    unsigned char *out = static_cast<unsigned char *>(out_results);
    const uint32_t par = h->pipelined ? static_cast<uint32_t>(h->call_no & 1u) : 0u;
    SegScratch &g = h->seg[par];
    SegScratch &prev = h->seg[h->pipelined ? (par ^ 1u) : 0u];

    std::vector<DevBatch> hb; // empty batches trigger nothing: only the non-empty ones reach the device
    hb.reserve(nbatches);
    uint64_t total = 0; uint32_t tiles = 0;
    uint64_t span_begin = ~0ull, span_end = 0; // address span of the segment's tuples (for the 2-D tensor map)
    for (uint32_t i = 0; i < nbatches; i++) {
        if (batches_h[i].n == 0) continue;
        if (!batches_h[i].tuples) return WFB_E_BADARG;
        DevBatch b; std::memset(&b, 0, sizeof(b));
        b.tuples = static_cast<const unsigned char *>(batches_h[i].tuples); b.ts = batches_h[i].ts;
        b.watermark = batches_h[i].watermark; b.n = batches_h[i].n; b.tile_begin = tiles;
        tiles += tiles_of(b.n); total += b.n;
        const uint64_t p0 = reinterpret_cast<uint64_t>(b.tuples);
        span_begin = std::min(span_begin, p0); span_end = std::max(span_end, p0 + static_cast<uint64_t>(b.n) * h->ops->tuple_bytes);
        hb.push_back(b);
    }
    if (total > 0x7fffffffull) return WFB_E_BADARG;
    if (total == 0) { // nothing to ingest: (pipelined) still deliver what is pending
        if (h->pipelined) return ffat_deliver(h, prev, out, out_ts, out_capacity, n_out_dev, s);
        if (!h->append_results) CK(cudaMemsetAsync(n_out_dev, 0, sizeof(uint32_t), s));
        return 0;
    }
    nbatches = static_cast<uint32_t>(hb.size());
    SEC(0);
    // bucket path: no global compaction in the streaming pass -- tile t owns positions [t*TILE, +TILE) of the segment
    const bool sparse = h->buckets && h->sparse_ingest;
    const uint64_t seg_cap = sparse ? static_cast<uint64_t>(tiles) * TILE : total;
    if (seg_cap > 0x7fffffffull) return WFB_E_BADARG;
    rc = ffat_ensure_segment(h, g, static_cast<uint32_t>(seg_cap), nbatches, s); if (rc) return rc;
    rc = h->ts.ensure_tiles(tiles); if (rc) return rc;
    { int rc_ = h->ts.stage.h2d(g.d_batches, hb.data(), sizeof(DevBatch) * nbatches, s); if (rc_) return rc_; }
    g.nbatches = nbatches; g.total = static_cast<uint32_t>(seg_cap); g.sparse = sparse;
    if (sparse) { // first position of every batch (the compacting pass writes the compact offsets itself)
        std::vector<uint32_t> boff(nbatches + 1);
        for (uint32_t i = 0; i < nbatches; i++) boff[i] = hb[i].tile_begin * TILE;
        boff[nbatches] = tiles * TILE;
        { int rc_ = h->ts.stage.h2d(g.batch_off, boff.data(), sizeof(uint32_t) * (nbatches + 1), s); if (rc_) return rc_; }
    }

    SEC(1);
    FfatDev ff = h->ff; // this call's view of the state: per-segment buffers of parity `par`
    ff.seg_cnt = g.seg_cnt; ff.trig = g.trig; ff.n_trig = g.n_trig; ff.trig_cap = g.trig_cap; ff.n_heavy = g.n_heavy;

    const uint32_t npasses = h->buckets ? 1u : h->sort_passes;
    const uint32_t pshift = h->buckets ? h->bucket_shift : 0u;
    const bool fuse_hist = npasses <= 4; // the streaming pass also counts the digits of the slot sort that follows
    if (fuse_hist) { rc = h->buckets ? RadixSorter::prepare_wide(g.sort_ctl, s) : RadixSorter::prepare(g.sort_ctl, npasses, s); if (rc) return rc; }
    h->mark(0, s);
    // 1. streaming pass: [map -> filter ->] lift, key -> slot, stable compaction over the whole segment
    TileArgs a; std::memset(&a, 0, sizeof(a));
    if (fuse_hist) { a.sort_ctl = g.sort_ctl; a.sort_passes = npasses; a.sort_shift = pshift; a.sort_dbits = h->buckets ? OSW_BITS : 8u; }
    g.hist_ready = fuse_hist;
    a.batches = g.d_batches; a.nbatches = nbatches; a.num_tiles = tiles;
    a.lifted = g.lifted; a.slots = g.slotsA; a.batch_off = g.batch_off; a.n_total = g.n_total; a.ff = ff;
    g.lifted_src = g.lifted;
    if (sparse && !h->pipelined && !h->bucket_move && (h->ops->reserved & 1u) && h->inplace_ok) {
        // pass-through program and every batch at its tile position inside one buffer: read the records where they are
        const unsigned char *base = hb[0].tuples;
        bool ok = (reinterpret_cast<uintptr_t>(base) & 15u) == 0;
        for (uint32_t i = 0; ok && i < nbatches; i++) ok = hb[i].tuples == base + static_cast<size_t>(hb[i].tile_begin) * TILE * h->ops->tuple_bytes;
        if (ok) { a.inplace = 1; g.lifted_src = base; a.ext_slots = ext_slots; }
    }
    if (ext_slots != nullptr && !a.inplace) return WFB_E_UNSUPPORTED; // (the front end always meets the in-place conditions)
    h->ts.next_launch(a);
    a.max_ctas_per_sm = h->ingest_ctas_per_sm;
    a.l2_hints = h->l2_hints ? 1u : 0u;
    a.sparse = sparse ? 1u : 0u;
    a.count_keys = h->buckets ? 0u : 1u;
    g.h32_ready = false;
    if (sparse && fuse_hist && h->fuse_tile_hist && !h->pipelined) { // (pipelined: the rows would be shared by two segments in flight)
        // the tile pass also files its digit counts per tile of the wide partition
        rc = h->sorter.prepare_h32(g.total, s, &a.wide_h32); if (rc) return rc;
        g.h32_ready = true;
    }
    g.h16_ready = false; g.ranked = false;
    if (a.inplace && a.sort_passes <= 1 && h->ops->slots_inplace && h->inplace_kernel) {
        // records read in place: only the slots (and the digit counts) are produced -- no tiles to stage, a plain kernel does it
        if (sparse && fuse_hist && h->tile_h16 && !h->pipelined && !g.h32_ready) { // whole wide tiles per CTA: rows + ranks for the partition, as the tile pass does
            rc = h->sorter.ensure_wide(g.total, s, &a.wide_h16); if (rc) return rc;
            a.sort_ctl = nullptr;
            g.h16_ready = true;
            a.pack_rank = (h->rank_scatter && ff.max_keys <= 65536u) ? 1u : 0u;
            g.ranked = a.pack_rank != 0;
        }
        rc = h->ops->slots_inplace(a, pre ? static_cast<const void *>(pre) : h->pp(), s); if (rc) return rc;
    } else {
        uint32_t claims = tiles;
        if (sparse && fuse_hist && h->tile_h16 && !g.h32_ready) {
            // a CTA claims the 16 tiles of a wide tile at once, counts its digits in shared memory and files the row itself:
            // the partition that follows needs neither a counting pass nor the per-CTA global digit counts
            if (h->pipelined && (g.total + OSW_TILE - 1) / OSW_TILE > h->sorter.wide_tiles) CK(cudaStreamSynchronize(h->s2)); // (the rows are about to be re-allocated)
            rc = h->sorter.ensure_wide(g.total, s, &a.wide_h16); if (rc) return rc; // (also sizes the chunk rows the partition needs)
            if (h->pipelined) {
                const uint32_t wt = (g.total + OSW_TILE - 1) / OSW_TILE;
                if (wt > g.h16_tiles) {
                    CK(cudaStreamSynchronize(s)); CK(cudaStreamSynchronize(h->s2));
                    cudaFree(g.h16);
                    g.h16_tiles = std::max(wt, 2 * g.h16_tiles);
                    CK(cudaMalloc(&g.h16, sizeof(uint16_t) * OSW_DIGITS * g.h16_tiles));
                }
                a.wide_h16 = g.h16;
            }
            a.tiles_per_ticket = OSW_TILE_POS / TILE;
            a.sort_ctl = nullptr; // (the chunk-sum kernel accumulates the global counts into g.sort_ctl, cleared above)
            claims = (tiles + a.tiles_per_ticket - 1) / a.tiles_per_ticket;
            g.h16_ready = true;
            a.pack_rank = (h->rank_scatter && ff.max_keys <= 65536u) ? 1u : 0u;
            g.ranked = a.pack_rank != 0;
        }
        uint32_t grid = 0;
        rc = h->ops->tile_pass(MODE_INGEST, a, pre ? static_cast<const void *>(pre) : h->pp(), claims, s, &grid, span_begin, span_end); if (rc) return rc;
        h->ts.launched(claims, grid);
    }
    h->launches++;
    h->mark(1, s);
    SEC(2);

    if (!h->pipelined) {
        if (!h->append_results) CK(cudaMemsetAsync(n_out_dev, 0, sizeof(uint32_t), s));
        rc = ffat_window_phase(h, g, ff, out, out_ts, out_capacity, n_out_dev, s); if (rc) return rc;
        h->mark(3, s);
        SEC(5);
    } else {
        // the ingest pass of this segment is queued: now hand over the previous segment's results, then start this
        // segment's sort + update on the internal stream, where it overlaps the NEXT call's ingest pass
        CK(cudaEventRecord(g.ev_ingest, s));
        rc = ffat_deliver(h, prev, out, out_ts, out_capacity, n_out_dev, s); if (rc) return rc;
        CK(cudaStreamWaitEvent(h->s2, g.ev_ingest, 0));
        CK(cudaMemsetAsync(g.res_n, 0, sizeof(uint32_t), h->s2));
        rc = ffat_window_phase(h, g, ff, g.res, g.res_ts, g.res_cap, g.res_n, h->s2); if (rc) return rc;
        h->mark(3, h->s2);
        CK(cudaEventRecord(g.ev_done, h->s2));
        g.pending = true;
    }
    if (h->timing && h->tev_used < wfb_ffat::TEV_MAX) h->tev_used++;
    h->call_no++;
    return 0;
}

// Window update on records that arrive grouped by bucket (destination side of the bucketed multi-GPU exchange): source s delivered
// offs_h[s+1] - offs_h[s] records from position offs_h[s] of `records`, bucket after bucket of this handle's slot space (buckets of
// 2^shift slots, bps of them; bins = the run lengths [nsrc][bps], recv_slots = the records' slots before masking). No partition pass:
// the runs of a bucket, in source order, ARE its items in stream order.
static int ffat_process_prebucketed(wfb_ffat *h, const unsigned char *records, const uint32_t *recv_slots, const uint32_t *bins, uint32_t nsrc,
                                    uint32_t bps, const uint32_t *offs_h, const uint64_t *wms_h, uint32_t slot_mask, uint32_t shift,
                                    void *out_results, uint64_t *out_ts, uint32_t out_capacity, uint32_t *n_out_dev, cudaStream_t s, bool append, uint32_t items)
{   // items: records delivered in all (the runs of a source need not be back to back with the next source's: offs_h[nsrc] only bounds the positions)
    if (!h || !n_out_dev || nsrc == 0 || nsrc > MAX_SHARDS || bps == 0 || bps > OSW_DIGITS || (1u << shift) > BK_KEYS) return WFB_E_BADARG;
    if (h->win_type != 0 || h->pipelined || !h->buckets) return WFB_E_UNSUPPORTED;
    int rc = h->ts.enter(s); if (rc) return rc;
    SegScratch &g = h->seg[0];
    const uint32_t total = items;
    if (!append) CK(cudaMemsetAsync(n_out_dev, 0, sizeof(uint32_t), s)); // (append: the results follow the ones already in the buffer)
    if (total == 0) return 0;
    rc = ffat_ensure_segment(h, g, total, nsrc, s); if (rc) return rc;
    std::vector<DevBatch> hb(nsrc);
    for (uint32_t i = 0; i < nsrc; i++) {
        std::memset(&hb[i], 0, sizeof(DevBatch));
        hb[i].tuples = records + static_cast<size_t>(offs_h[i]) * h->ops->result_bytes; hb[i].n = offs_h[i + 1] - offs_h[i]; hb[i].watermark = wms_h[i];
    }
    { int rc_ = h->ts.stage.h2d(g.d_batches, hb.data(), sizeof(DevBatch) * nsrc, s); if (rc_) return rc_; }
    { int rc_ = h->ts.stage.h2d(g.batch_off, offs_h, sizeof(uint32_t) * (nsrc + 1), s); if (rc_) return rc_; }
    g.nbatches = nsrc; g.total = total; g.sparse = true; g.lifted_src = records;
    FfatDev ff = h->ff;
    ff.seg_cnt = g.seg_cnt; ff.trig = g.trig; ff.n_trig = g.n_trig; ff.trig_cap = g.trig_cap; ff.n_heavy = g.n_heavy;
    rc = RadixSorter::prepare_wide(g.sort_ctl, s); if (rc) return rc; // (buckets at or above bps stay empty)
    h->mark(0, s); h->mark(1, s);
    MgRuns runs;
    for (uint32_t i = 0; i <= MAX_SHARDS; i++) runs.off[i] = offs_h[std::min(i, nsrc)];
    // the sources' 1024 bins are shared by all destinations: split every coarse bucket so that the update kernel gets (up to) 1024 buckets
    uint32_t nsub = 1, shift2 = shift;
    while (nsub * 2 * bps <= OSW_DIGITS && nsub * 2 <= MAX_SHARDS && shift2 > 0) { nsub *= 2; shift2--; }
    if (!h->mg_scratch) CK(cudaMalloc(&h->mg_scratch, sizeof(uint32_t) * 3 * OSW_DIGITS * MAX_SHARDS));
    uint32_t *cnt3 = h->mg_scratch, *off3 = cnt3 + OSW_DIGITS * MAX_SHARDS, *run_starts = off3 + OSW_DIGITS * MAX_SHARDS;
    const dim3 grid(bps, nsrc);
    k_mg_count<<<grid, MG_THREADS, 0, s>>>(bins, nsrc, bps, runs, recv_slots, slot_mask, shift2, nsub, cnt3, run_starts, g.n_trig, g.n_heavy);
    k_mg_scan<<<1, 1024, 0, s>>>(cnt3, bps * nsub * nsrc, nsrc, off3, g.sort_ctl);
    k_mg_split<<<grid, MG_THREADS, 0, s>>>(bins, nsrc, bps, runs, recv_slots, slot_mask, shift2, nsub, off3, run_starts, g.slotsB, g.posB);
    CK(cudaGetLastError());
    h->mark(2, s);
    rc = h->ops->ffat_buckets(ff, records, g.slotsB, g.posB, g.sort_ctl, shift2, 0u, g.batch_off, g.d_batches, nsrc, static_cast<unsigned char *>(out_results), out_ts,
                              out_capacity, n_out_dev, s, h->pp());
    if (rc) return rc;
    rc = h->ops->ffat_windows(ff, g.batch_off, g.d_batches, nsrc, static_cast<unsigned char *>(out_results), out_ts, out_capacity, static_cast<uint32_t>(g_num_sms) * 4u, s,
                              h->pp(), n_out_dev);
    if (rc) return rc;
    h->mark(3, s);
    h->launches += 5;
    if (h->timing && h->tev_used < wfb_ffat::TEV_MAX) h->tev_used++;
    h->call_no++;
    return 0;
}

int wfb_ffat_flush(wfb_ffat_t *h, void *out_results, uint64_t *out_ts, uint32_t out_capacity, uint32_t *n_out_dev, void *stream)
{
    if (!h || !n_out_dev || (out_capacity && !out_results)) return WFB_E_BADARG;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (!h->pipelined) { CK(cudaMemsetAsync(n_out_dev, 0, sizeof(uint32_t), s)); return 0; }
    int rc = h->ts.enter(s); if (rc) return rc;
    // at most one segment is pending: the one of the last call
    SegScratch &last = h->seg[(h->call_no + 1) & 1u];
    return ffat_deliver(h, last, static_cast<unsigned char *>(out_results), out_ts, out_capacity, n_out_dev, s);
}

int wfb_ffat_timing(wfb_ffat_t *h, int enable, float *ms_h, uint32_t *calls_h)
{
    if (!h) return WFB_E_BADARG;
    float acc[4] = {0, 0, 0, 0};
    if (h->tev_used) {
        CK(cudaEventSynchronize(h->tev[(h->tev_used - 1) * 4 + 3]));
        for (uint32_t i = 0; i < h->tev_used; i++) {
            float a = 0, b = 0, c = 0, d = 0;
            cudaEvent_t *e = &h->tev[i * 4];
            CK(cudaEventElapsedTime(&a, e[0], e[1]));
            CK(cudaEventElapsedTime(&b, e[1], e[2]));
            CK(cudaEventElapsedTime(&c, e[2], e[3]));
            CK(cudaEventElapsedTime(&d, e[0], e[3]));
            acc[0] += a; acc[1] += b; acc[2] += c; acc[3] += d;
        }
    }
    if (ms_h) for (int i = 0; i < 4; i++) ms_h[i] = acc[i];
    if (calls_h) *calls_h = h->tev_used;
    h->tev_used = 0;
    if (enable && h->tev.empty()) {
        h->tev.resize(wfb_ffat::TEV_MAX * 4);
        for (auto &e : h->tev) CK(cudaEventCreate(&e));
    }
    h->timing = enable != 0;
    return 0;
}

int wfb_ffat_stats(wfb_ffat_t *h, uint32_t *n_keys_h, uint32_t *err_flags_h, void *stream)
{
    if (!h) return WFB_E_BADARG;
    uint32_t v[2] = {0, 0};
    if (h->s2) CK(cudaStreamSynchronize(h->s2));
    CK(cudaMemcpyAsync(v, h->ff.n_slots, sizeof(v), cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(stream)));
    CK(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
    if (n_keys_h) *n_keys_h = v[0];
    if (err_flags_h) *err_flags_h = v[1];
    if (h->cb && err_flags_h) { uint32_t e2 = 0; int rc = wfb_ffat_stats(h->cb, nullptr, &e2, stream); if (rc) return rc; *err_flags_h |= e2; }
    return 0;
}

int wfb_ffat_results_total(wfb_ffat_t *h, uint64_t *total_h, void *stream)
{
    if (!h || !total_h) return WFB_E_BADARG;
    const wfb_ffat *src = h->cb ? h->cb : h; // time-based handles: the count-based back end emits the results
    unsigned long long v = 0;
    if (src->s2) CK(cudaStreamSynchronize(src->s2));
    if (src->ff.results_total == nullptr) { *total_h = 0; return 0; }
    CK(cudaMemcpyAsync(&v, src->ff.results_total, sizeof(v), cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(stream)));
    CK(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
    *total_h = v;
    return 0;
}

// ---- key-sharded pipeline across GPUs ------------------------------------------------------------------------------------------------
} // extern "C"
#include <dlfcn.h>
namespace {
// the few NCCL entry points used, resolved at run time (the library a torch process has already loaded, or the system one)
struct Nccl {
    struct Id { char b[128]; }; // ncclUniqueId (passed by value)
    typedef int (*GetUniqueId_t)(void *);
    typedef int (*CommInitRank_t)(void **, int, Id, int);
    typedef int (*CommDestroy_t)(void *);
    typedef int (*SendRecv_t)(void *, size_t, int, int, void *, cudaStream_t);
    typedef int (*Group_t)();
    typedef const char *(*ErrStr_t)(int);
    void *lib = nullptr;
    GetUniqueId_t GetUniqueId = nullptr; CommInitRank_t CommInitRank = nullptr; CommDestroy_t CommDestroy = nullptr;
    SendRecv_t Send = nullptr, Recv = nullptr; Group_t GroupStart = nullptr, GroupEnd = nullptr;
    bool ok = false;
    Nccl()
    {
        const char *names[] = {std::getenv("WFB_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char *n : names) { if (n && (lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break; }
        if (!lib) return;
        GetUniqueId = reinterpret_cast<GetUniqueId_t>(dlsym(lib, "ncclGetUniqueId"));
        CommInitRank = reinterpret_cast<CommInitRank_t>(dlsym(lib, "ncclCommInitRank"));
        CommDestroy = reinterpret_cast<CommDestroy_t>(dlsym(lib, "ncclCommDestroy"));
        Send = reinterpret_cast<SendRecv_t>(dlsym(lib, "ncclSend")); Recv = reinterpret_cast<SendRecv_t>(dlsym(lib, "ncclRecv"));
        GroupStart = reinterpret_cast<Group_t>(dlsym(lib, "ncclGroupStart")); GroupEnd = reinterpret_cast<Group_t>(dlsym(lib, "ncclGroupEnd"));
        ok = GetUniqueId && CommInitRank && CommDestroy && Send && Recv && GroupStart && GroupEnd;
    }
};
Nccl &nccl() { static Nccl n; return n; }
constexpr int NCCL_UINT8 = 1; // ncclUint8
#define NK(call) do { int r__ = (call); if (r__ != 0) return 1000 + r__; } while (0) // (NCCL errors: 1000 + ncclResult_t)

__global__ void k_mg_meta(const uint32_t *__restrict__ counts, uint64_t watermark, uint32_t nranks, uint64_t *__restrict__ send_meta)
{
    const uint32_t d = threadIdx.x;
    if (d < nranks) { send_meta[2 * d] = counts[d]; send_meta[2 * d + 1] = watermark; }
}

// appended results (wfb_mg_flush): the call about to run adds the WHOLE count of the output buffer to the handle's total
__global__ void k_mg_pre_append(unsigned long long *results_total, const uint32_t *n_out) { if (results_total) *results_total -= *n_out; }

struct MgSlot { // buffers of one step in flight (three: the exchange of step i-2 overlaps the source pass of step i)
    unsigned char *regions = nullptr; uint32_t region_cap = 0; // records by destination (bucketed: bin after bin, region_cap = the segment's positions)
    uint32_t *vslots = nullptr;                // bucketed: virtual slot of every record of `regions`
    uint32_t *bins = nullptr;                  // bucketed: OSW_DIGITS + 1 words, the bin sizes of this step's partition
    uint32_t *recv_slots = nullptr, *recv_bins = nullptr; size_t recv_slots_cap = 0; // bucketed: what the sources delivered ([nranks][bps] run lengths)
    unsigned char *ce_buf = nullptr;           // copy-engine exchange: the receive buffers of this slot in ONE allocation other ranks map (cudaIpc):
                                               // [records n x cap][slots n x cap][run lengths n x bps], source s at stride cap
    unsigned char *peer[MAX_SHARDS] = {};      // every rank's ce_buf of this slot, mapped here
    uint32_t offs[MAX_SHARDS + 1] = {}; uint64_t wms[MAX_SHARDS] = {}; // where every source's records start in `recv` / their watermarks (host, set by the exchange)
    uint32_t *counts = nullptr;                // MAX_SHARDS + 1 (device)
    uint64_t *send_meta = nullptr, *recv_meta = nullptr; // [nranks][2] (device)
    uint32_t *h_counts = nullptr; uint64_t *h_recv = nullptr; // pinned copies
    unsigned char *recv = nullptr; size_t recv_bytes = 0;
    cudaEvent_t ev_src = nullptr, ev_meta = nullptr, ev_a2a = nullptr, ev_done = nullptr, ev_self = nullptr, ev_fork = nullptr;
    bool used = false, done_recorded = false, exchanged = false;
    cudaEvent_t tr[8] = {}; bool tr_valid = false; // WFB_MG_TRACE: source begin/end, update begin/end (caller's stream); exchange begin/end, sizes begin/end (communication stream)
};
} // namespace

struct wfb_mg {
    int nranks = 1, rank = 0;
    void *comm = nullptr;
    wfb_engine_t *eng = nullptr;
    wfb_ffat_t *ffat = nullptr;
    size_t rb = 0;
    MgSlot slot[4];            // four steps in flight: step i's source pass, step i-1 waiting, step i-2 travelling, step i-3 being updated
    cudaStream_t cs = nullptr; // communication stream
    cudaStream_t cs2 = nullptr; // the rank's own share of an exchange: device-to-device copies (copy engine), next to the NCCL group
    uint64_t step_no = 0;
    MgSlot *pend[3] = {nullptr, nullptr, nullptr}; int npend = 0; // steps not yet updated, oldest first (the oldest of three has been exchanged)
    std::vector<wfb_batch_t> chunks;
    // bucketed exchange: the source partitions by (destination, bucket of the destination's slot space), the destination only concatenates runs
    // copy-engine exchange (bucketed mode, all ranks on one node): records are PUSHED into the peers' receive buffers with plain
    // device-to-device copies over NVLink (no SM, no NCCL channel limit); the next step's size exchange is the completion signal
    bool ce = false, ce_tried = false; uint64_t ce_cap = 0, peer_cap[MAX_SHARDS] = {};
    unsigned char *ce_msg = nullptr; // device staging of the handle exchange
    cudaEvent_t ev_flush = nullptr;
    static constexpr int CE_STREAMS = 4;
    cudaStream_t ce_s[CE_STREAMS] = {}; cudaEvent_t ce_ev[CE_STREAMS] = {}; // peer copies of one exchange are spread over these (several copy engines)
    uint32_t *tok = nullptr;            // device words of the completion tokens: [0] sent, [1 + p] received from p
    cudaEvent_t last_done = nullptr;    // end of the most recently issued window update (caller's stream)
    double host_acc[3] = {}; uint64_t host_n = 0;
    bool trace = false; double tr_acc[8] = {}; uint64_t tr_n = 0; // WFB_MG_TRACE=1: device timeline of a step, printed every 64 steps (tuning aid)
    bool bucketed = false;
    uint32_t shard_slots = 0, shard_keys = 0, shift = 0, bps = 0; // slots per destination (power of two), keys per destination, bucket = slot >> shift, buckets per destination
};

extern "C" {

int wfb_mg_unique_id(void *id128_h)
{
    if (!id128_h) return WFB_E_BADARG;
    if (!nccl().ok) return WFB_E_UNSUPPORTED;
    NK(nccl().GetUniqueId(id128_h));
    return 0;
}

int wfb_mg_destroy(wfb_mg_t *h)
{
    if (!h) return 0;
    cudaDeviceSynchronize();
    if (h->comm && nccl().ok) nccl().CommDestroy(h->comm);
    if (h->eng) wfb_engine_destroy(h->eng);
    if (h->ffat) wfb_ffat_destroy(h->ffat);
    for (MgSlot &sl : h->slot) {
        cudaFree(sl.regions); cudaFree(sl.counts); cudaFree(sl.send_meta); cudaFree(sl.recv_meta); if (!sl.ce_buf) cudaFree(sl.recv);
        cudaFree(sl.vslots); cudaFree(sl.bins);
        if (sl.ce_buf) { // (recv / recv_slots / recv_bins point into ce_buf)
            for (int p = 0; p < h->nranks; p++) if (p != h->rank && sl.peer[p]) cudaIpcCloseMemHandle(sl.peer[p]);
            cudaFree(sl.ce_buf); sl.recv = nullptr;
        } else { cudaFree(sl.recv_slots); cudaFree(sl.recv_bins); }
        if (sl.h_counts) cudaFreeHost(sl.h_counts);
        if (sl.h_recv) cudaFreeHost(sl.h_recv);
        for (cudaEvent_t e : {sl.ev_src, sl.ev_meta, sl.ev_a2a, sl.ev_done, sl.ev_self, sl.ev_fork}) if (e) cudaEventDestroy(e);
        for (cudaEvent_t e : sl.tr) if (e) cudaEventDestroy(e);
    }
    if (h->cs) cudaStreamDestroy(h->cs);
    if (h->cs2) cudaStreamDestroy(h->cs2);
    cudaFree(h->ce_msg); cudaFree(h->tok);
    if (h->ev_flush) cudaEventDestroy(h->ev_flush);
    for (cudaStream_t st : h->ce_s) if (st) cudaStreamDestroy(st);
    for (cudaEvent_t e : h->ce_ev) if (e) cudaEventDestroy(e);
    delete h;
    cudaGetLastError();
    return 0;
}

int wfb_mg_create(wfb_mg_t **hh, int prog, int nranks, int rank, const void *id128_h, uint64_t win, uint64_t slide, uint32_t wins_per_batch,
                  uint32_t max_keys_total)
{
    if (!hh || nranks < 1 || nranks > static_cast<int>(MAX_SHARDS) || rank < 0 || rank >= nranks || (nranks > 1 && !id128_h) || max_keys_total == 0) return WFB_E_BADARG;
    const ProgramOps *o = program(prog);
    if (!o) return WFB_E_NOPROG;
    const int lp = lifted_program_of(prog);
    if (lp < 0 || !(program(lp)->reserved2 & 1u)) return WFB_E_UNSUPPORTED; // the lifted records must carry their key (Program::result_key)
    int rc = device_ready(); if (rc) return rc;
    if (nranks > 1 && !nccl().ok) return WFB_E_UNSUPPORTED;
    wfb_mg *h = new (std::nothrow) wfb_mg();
    if (!h) return WFB_E_BADARG;
    h->nranks = nranks; h->rank = rank; h->rb = o->result_bytes;
    h->trace = std::getenv("WFB_MG_TRACE") && std::atoi(std::getenv("WFB_MG_TRACE")) != 0;
#define MGCK(call) do { int r__ = (call); if (r__) { wfb_mg_destroy(h); return r__; } } while (0)
    MGCK(wfb_engine_create(&h->eng, prog));
    // the rank's replica owns the keys with key % nranks == rank: compact slots key / nranks, records read in place
    MGCK(wfb_ffat_create(&h->ffat, lp, win, slide, wins_per_batch, (max_keys_total + nranks - 1) / nranks, 0, 0, WFB_FFAT_DENSE_KEYS));
    if (nranks > 1) MGCK(wfb_ffat_set_key_shard(h->ffat, static_cast<uint32_t>(nranks), static_cast<uint32_t>(rank)));
    MGCK(static_cast<int>(cudaStreamCreateWithFlags(&h->cs, cudaStreamNonBlocking)));
    MGCK(static_cast<int>(cudaStreamCreateWithFlags(&h->cs2, cudaStreamNonBlocking)));
    MGCK(static_cast<int>(cudaEventCreateWithFlags(&h->ev_flush, cudaEventDisableTiming)));
    {   // bucketed exchange when the destination-major virtual slots fit 16 bits (they travel packed with a 16-bit rank)
        const uint32_t keys = (max_keys_total + nranks - 1) / nranks;
        uint32_t L = 1; while (L < keys) L <<= 1;
        static const bool off = std::getenv("WFB_MG_BUCKETED") && std::atoi(std::getenv("WFB_MG_BUCKETED")) == 0;
        const size_t rb = o->result_bytes;
        if (!off && static_cast<uint64_t>(L) * nranks <= 65536u && h->ffat->buckets && (rb == 16 || rb == 24 || rb == 32 || rb == 48 || rb == 64)) {
            uint32_t span = 1; while (span < L * static_cast<uint32_t>(nranks)) span <<= 1; // virtual slot space, rounded up
            uint32_t sh = 0; while ((span >> sh) > OSW_DIGITS) sh++;
            if ((L >> sh) >= 1) { h->bucketed = true; h->shard_slots = L; h->shard_keys = keys; h->shift = sh; h->bps = L >> sh; }
        }
    }
    for (MgSlot &sl : h->slot) {
        if (h->bucketed) {
            MGCK(static_cast<int>(cudaMalloc(&sl.bins, sizeof(uint32_t) * (OSW_DIGITS + 1))));
            MGCK(static_cast<int>(cudaMalloc(&sl.recv_bins, sizeof(uint32_t) * MAX_SHARDS * OSW_DIGITS)));
        }
        MGCK(static_cast<int>(cudaMalloc(&sl.counts, sizeof(uint32_t) * (MAX_SHARDS + 1))));
        MGCK(static_cast<int>(cudaMalloc(&sl.send_meta, sizeof(uint64_t) * 2 * MAX_SHARDS)));
        MGCK(static_cast<int>(cudaMalloc(&sl.recv_meta, sizeof(uint64_t) * 2 * MAX_SHARDS)));
        MGCK(static_cast<int>(cudaMallocHost(&sl.h_counts, sizeof(uint32_t) * (MAX_SHARDS + 1))));
        MGCK(static_cast<int>(cudaMallocHost(&sl.h_recv, sizeof(uint64_t) * 2 * MAX_SHARDS)));
        for (cudaEvent_t *e : {&sl.ev_src, &sl.ev_meta, &sl.ev_a2a, &sl.ev_done, &sl.ev_self, &sl.ev_fork}) MGCK(static_cast<int>(cudaEventCreateWithFlags(e, cudaEventDisableTiming)));
        if (h->trace) for (cudaEvent_t &e : sl.tr) MGCK(static_cast<int>(cudaEventCreate(&e)));
    }
    if (nranks > 1) {
        Nccl::Id id; std::memcpy(id.b, id128_h, sizeof(id.b));
        int r = nccl().CommInitRank(&h->comm, nranks, id, rank);
        if (r != 0) { wfb_mg_destroy(h); return 1000 + r; }
    }
#undef MGCK
    *hh = h;
    return 0;
}

// ---- copy-engine exchange: setup (collective, at the first step) ---------------------------------------------------------------
struct CeLayout { size_t slots_off, bins_off, bytes; };
static CeLayout ce_layout(uint64_t cap, int n, size_t rb, uint32_t bps)
{
    CeLayout l; l.slots_off = static_cast<size_t>(n) * cap * rb; l.bins_off = (l.slots_off + static_cast<size_t>(n) * cap * 4 + 255) & ~static_cast<size_t>(255);
    l.bytes = l.bins_off + static_cast<size_t>(n) * bps * 4; return l;
}
constexpr int MG_SLOTS = 4;
struct CeMsg { cudaIpcMemHandle_t h[MG_SLOTS]; uint64_t cap; uint64_t ok; };
// one small message to / from every other rank (communication stream, synchronous)
static int mg_all_exchange(wfb_mg *h, const void *mine_h, void *all_h, size_t bytes)
{
    const int n = h->nranks;
    if (!h->ce_msg) CK(cudaMalloc(&h->ce_msg, sizeof(CeMsg) * (MAX_SHARDS + 1)));
    if (bytes > sizeof(CeMsg)) return WFB_E_BADARG;
    unsigned char *snd = h->ce_msg, *rcv = h->ce_msg + sizeof(CeMsg);
    CK(cudaMemcpyAsync(snd, mine_h, bytes, cudaMemcpyHostToDevice, h->cs));
    NK(nccl().GroupStart());
    for (int p = 0; p < n; p++) {
        if (p == h->rank) continue;
        NK(nccl().Send(snd, bytes, NCCL_UINT8, p, h->comm, h->cs));
        NK(nccl().Recv(rcv + static_cast<size_t>(p) * bytes, bytes, NCCL_UINT8, p, h->comm, h->cs));
    }
    NK(nccl().GroupEnd());
    CK(cudaMemcpyAsync(all_h, rcv, bytes * n, cudaMemcpyDeviceToHost, h->cs));
    CK(cudaStreamSynchronize(h->cs));
    std::memcpy(static_cast<unsigned char *>(all_h) + static_cast<size_t>(h->rank) * bytes, mine_h, bytes);
    return 0;
}
static int mg_ce_setup(wfb_mg *h, uint64_t positions)
{
    h->ce_tried = true;
    static const bool off = std::getenv("WFB_MG_CE") && std::atoi(std::getenv("WFB_MG_CE")) == 0;
    if (off || !h->bucketed || h->nranks < 2) return 0;
    const int n = h->nranks;
    const uint64_t cap = (positions + 1023) & ~1023ull; // worst case: every survivor of a source's step goes to one destination
    const CeLayout l = ce_layout(cap, n, h->rb, h->bps);
    CeMsg mine; std::memset(&mine, 0, sizeof(mine)); mine.cap = cap; mine.ok = 1;
    for (int i = 0; i < MG_SLOTS && mine.ok; i++) {
        if (cudaMalloc(&h->slot[i].ce_buf, l.bytes) != cudaSuccess || cudaIpcGetMemHandle(&mine.h[i], h->slot[i].ce_buf) != cudaSuccess) mine.ok = 0;
    }
    cudaGetLastError();
    CeMsg all[MAX_SHARDS];
    int rc = mg_all_exchange(h, &mine, all, sizeof(CeMsg)); if (rc) return rc;
    uint64_t ok = 1;
    for (int p = 0; p < n; p++) ok &= all[p].ok;
    if (ok) {
        for (int p = 0; p < n && ok; p++) {
            h->peer_cap[p] = all[p].cap;
            for (int i = 0; i < MG_SLOTS && ok; i++) {
                if (p == h->rank) { h->slot[i].peer[p] = h->slot[i].ce_buf; continue; }
                void *ptr = nullptr;
                if (cudaIpcOpenMemHandle(&ptr, all[p].h[i], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; cudaGetLastError(); }
                else h->slot[i].peer[p] = static_cast<unsigned char *>(ptr);
            }
        }
    }
    // second round: every rank could map every buffer, or nobody uses the path
    uint64_t st_mine = ok, st_all[MAX_SHARDS];
    rc = mg_all_exchange(h, &st_mine, st_all, sizeof(uint64_t)); if (rc) return rc;
    for (int p = 0; p < n; p++) ok &= st_all[p];
    if (!ok) {
        for (MgSlot &sl : h->slot) {
            for (int p = 0; p < n; p++) { if (p != h->rank && sl.peer[p]) cudaIpcCloseMemHandle(sl.peer[p]); sl.peer[p] = nullptr; }
            cudaFree(sl.ce_buf); sl.ce_buf = nullptr;
        }
        cudaGetLastError();
        if (h->trace) std::fprintf(stderr, "[wfb_mg rank %d] copy-engine exchange not available (cudaIpc): NCCL exchange\n", h->rank);
        return 0; // (the NCCL exchange stays in use)
    }
    for (int i = 0; i < wfb_mg::CE_STREAMS; i++) { CK(cudaStreamCreateWithFlags(&h->ce_s[i], cudaStreamNonBlocking)); CK(cudaEventCreateWithFlags(&h->ce_ev[i], cudaEventDisableTiming)); }
    CK(cudaMalloc(&h->tok, sizeof(uint32_t) * (MAX_SHARDS + 1)));
    CK(cudaMemset(h->tok, 0, sizeof(uint32_t) * (MAX_SHARDS + 1)));
    h->ce = true; h->ce_cap = cap;
    if (h->trace) std::fprintf(stderr, "[wfb_mg rank %d] copy-engine exchange enabled (capacity %llu records per source)\n", h->rank, static_cast<unsigned long long>(cap));
    for (MgSlot &sl : h->slot) {
        cudaFree(sl.recv); cudaFree(sl.recv_slots); cudaFree(sl.recv_bins);
        sl.recv = sl.ce_buf; sl.recv_slots = reinterpret_cast<uint32_t *>(sl.ce_buf + l.slots_off); sl.recv_bins = reinterpret_cast<uint32_t *>(sl.ce_buf + l.bins_off);
        sl.recv_bytes = l.slots_off; sl.recv_slots_cap = static_cast<size_t>(n) * cap;
    }
    return 0;
}

// source side of a step: fused pass + partition by destination; the sizes travel (and reach the host a step later)
static int mg_source(wfb_mg *h, MgSlot &sl, const wfb_functors_t *pre, const wfb_batch_t *batches_h, uint32_t nbatches, uint64_t watermark, cudaStream_t s)
{
    uint64_t n = 0;
    for (uint32_t i = 0; i < nbatches; i++) n += batches_h[i].n;
    if (n > 0x7fffffffull) return WFB_E_BADARG;
    int rc;
    if (h->trace) {
        if (sl.tr_valid) { // this slot's previous step (three calls ago) is complete: add its timeline
            CK(cudaEventSynchronize(sl.tr[3]));
            float ms; const int pairs[6][2] = {{0, 1}, {2, 3}, {4, 5}, {6, 7}, {4, 2}, {5, 2}};
            for (int i = 0; i < 6; i++) if (cudaEventElapsedTime(&ms, sl.tr[pairs[i][0]], sl.tr[pairs[i][1]]) == cudaSuccess) h->tr_acc[i] += ms;
            if (++h->tr_n % 64 == 0) {
                std::fprintf(stderr, "[wfb_mg rank %d] us/step over 64 steps: source %.0f | update %.0f | exchange %.0f | sizes %.0f | exchange begin -> update begin %.0f | exchange end -> update begin %.0f\n",
                             h->rank, h->tr_acc[0] / 64 * 1e3, h->tr_acc[1] / 64 * 1e3, h->tr_acc[2] / 64 * 1e3, h->tr_acc[3] / 64 * 1e3, h->tr_acc[4] / 64 * 1e3, h->tr_acc[5] / 64 * 1e3);
                for (double &a : h->tr_acc) a = 0;
            }
            sl.tr_valid = false;
        }
        CK(cudaEventRecord(sl.tr[0], s));
    }
    if (h->bucketed) {
        uint64_t positions = 0;
        for (uint32_t i = 0; i < nbatches; i++) positions += static_cast<uint64_t>(tiles_of(batches_h[i].n)) * TILE;
        if (!h->ce_tried) { rc = mg_ce_setup(h, std::max<uint64_t>(positions, 1)); if (rc) return rc; } // (collective: every rank is in its first step)
        if (sl.region_cap < positions) {
            if (sl.used) CK(cudaDeviceSynchronize());
            cudaFree(sl.regions); cudaFree(sl.vslots);
            sl.region_cap = static_cast<uint32_t>(positions);
            CK(cudaMalloc(&sl.regions, static_cast<size_t>(sl.region_cap) * h->rb));
            CK(cudaMalloc(&sl.vslots, sizeof(uint32_t) * sl.region_cap));
        }
        rc = shard_lift_impl(h->eng, pre, batches_h, nbatches, static_cast<uint32_t>(h->nranks), sl.regions, sl.region_cap, sl.counts, s,
                             h->shard_slots, h->shard_keys, h->shift, sl.vslots, sl.bins, n ? sl.send_meta : nullptr, watermark);
    } else {
        if (sl.region_cap < n) { // worst case: every item of the segment survives and goes to one shard
            if (sl.used) CK(cudaDeviceSynchronize());
            cudaFree(sl.regions);
            sl.region_cap = static_cast<uint32_t>(n);
            CK(cudaMalloc(&sl.regions, static_cast<size_t>(h->nranks) * sl.region_cap * h->rb));
        }
        rc = wfb_shard_lift(h->eng, pre, batches_h, nbatches, static_cast<uint32_t>(h->nranks), sl.regions, sl.region_cap, sl.counts, s);
    }
    if (rc) return rc;
    if (!(h->bucketed && n)) { // (bucketed: the kernel that sums the bins per destination wrote the pairs)
        k_mg_meta<<<1, 32, 0, s>>>(sl.counts, watermark, static_cast<uint32_t>(h->nranks), sl.send_meta);
        CK(cudaGetLastError());
    }
    CK(cudaEventRecord(sl.ev_src, s));
    if (h->trace) CK(cudaEventRecord(sl.tr[1], s));
    CK(cudaStreamWaitEvent(h->cs, sl.ev_src, 0));
    if (h->trace) CK(cudaEventRecord(sl.tr[6], h->cs));
    if (h->nranks > 1) {
        NK(nccl().GroupStart());
        for (int p = 0; p < h->nranks; p++) {
            NK(nccl().Send(sl.send_meta + 2 * p, 16, NCCL_UINT8, p, h->comm, h->cs));
            NK(nccl().Recv(sl.recv_meta + 2 * p, 16, NCCL_UINT8, p, h->comm, h->cs));
        }
        NK(nccl().GroupEnd());
    } else CK(cudaMemcpyAsync(sl.recv_meta, sl.send_meta, 16, cudaMemcpyDeviceToDevice, h->cs));
    CK(cudaMemcpyAsync(sl.h_counts, sl.counts, sizeof(uint32_t) * (MAX_SHARDS + 1), cudaMemcpyDeviceToHost, h->cs));
    CK(cudaMemcpyAsync(sl.h_recv, sl.recv_meta, sizeof(uint64_t) * 2 * h->nranks, cudaMemcpyDeviceToHost, h->cs));
    CK(cudaEventRecord(sl.ev_meta, h->cs));
    if (h->trace) CK(cudaEventRecord(sl.tr[7], h->cs));
    sl.used = true;
    return 0;
}

// exchange of the records of a step, on the communication stream (issued BEFORE the next step's source pass so that it runs next to it)
static int mg_exchange(wfb_mg *h, MgSlot &sl)
{
    CK(cudaEventSynchronize(sl.ev_meta)); // the sizes of this step on the host (a step old: no stall)
    if (sl.h_counts[MAX_SHARDS]) return WFB_E_CAPACITY; // a shard region overflowed / a key outside the declared key space
    const int n = h->nranks;
    if (h->bucketed) {
        // records, their slots and the run lengths of every source; source-rank order = global stream order
        uint64_t tot = 0;
        for (int p = 0; p < n; p++) { sl.offs[p] = static_cast<uint32_t>(tot); tot += sl.h_recv[2 * p]; sl.wms[p] = sl.h_recv[2 * p + 1]; }
        if (tot > 0x7fffffffull) return WFB_E_CAPACITY;
        sl.offs[n] = static_cast<uint32_t>(tot);
        if (h->ce) {
            // push: this rank's records for peer p go straight into p's receive buffer of this slot, at the stride-cap region of source `rank`
            const uint64_t cap = h->ce_cap;
            if (static_cast<uint64_t>(n) * cap > 0x7fffffffull) return WFB_E_CAPACITY;
            for (int p = 0; p <= n; p++) sl.offs[p] = static_cast<uint32_t>(p * cap);
            size_t send_off[MAX_SHARDS + 1]; send_off[0] = 0;
            for (int p = 0; p < n; p++) {
                send_off[p + 1] = send_off[p] + sl.h_counts[p];
                if (sl.h_recv[2 * p] > cap || sl.h_counts[p] > h->peer_cap[p]) return WFB_E_CAPACITY; // (a step larger than the first one: WFB_MG_CE=0)
            }
            const int slot_idx = static_cast<int>(&sl - h->slot);
            const size_t bin_bytes = sizeof(uint32_t) * h->bps;
            if (h->trace) CK(cudaEventRecord(sl.tr[4], h->cs));
            CK(cudaStreamWaitEvent(h->cs2, sl.ev_src, 0));
            if (sl.done_recorded) CK(cudaStreamWaitEvent(h->cs2, sl.ev_done, 0));
            // the peers' shares: plain copies (copy engines, no SM; 450 GB/s to one peer, ~250 GB/s aggregate with 7 peers x 3 pieces each), or --
            // WFB_MG_PUSH=sm -- one kernel that stores into the mapped buffers (16 CTAs per peer; 205 instead of 465 us at N = 8, same step time:
            // the exchange is hidden behind the source pass either way)
            static const char *push_env = std::getenv("WFB_MG_PUSH");
            const bool use_sm = push_env && std::strcmp(push_env, "sm") == 0 && h->rb % 16 == 0;
            {   // own share: a local copy next to the rest
                const int p = h->rank;
                const CeLayout l = ce_layout(h->peer_cap[p], n, h->rb, h->bps);
                unsigned char *dst = h->slot[slot_idx].peer[p];
                const uint64_t me = static_cast<uint64_t>(h->rank), pc = h->peer_cap[p];
                CK(cudaMemcpyAsync(dst + me * pc * h->rb, sl.regions + send_off[p] * h->rb, static_cast<size_t>(sl.h_counts[p]) * h->rb, cudaMemcpyDeviceToDevice, h->cs2));
                CK(cudaMemcpyAsync(dst + l.slots_off + me * pc * 4, sl.vslots + send_off[p], static_cast<size_t>(sl.h_counts[p]) * 4, cudaMemcpyDeviceToDevice, h->cs2));
                CK(cudaMemcpyAsync(dst + l.bins_off + me * bin_bytes, sl.bins + static_cast<size_t>(p) * h->bps, bin_bytes, cudaMemcpyDeviceToDevice, h->cs2));
                CK(cudaEventRecord(sl.ev_self, h->cs2));
            }
            if (use_sm) {
                MgPush a; std::memset(&a, 0, sizeof(a));
                int k = 0;
                for (int p = 0; p < n; p++) {
                    if (p == h->rank) continue;
                    const CeLayout l = ce_layout(h->peer_cap[p], n, h->rb, h->bps);
                    unsigned char *dst = h->slot[slot_idx].peer[p];
                    const uint64_t me = static_cast<uint64_t>(h->rank), pc = h->peer_cap[p];
                    a.rec_src[k] = reinterpret_cast<const uint4 *>(sl.regions + send_off[p] * h->rb); a.rec_dst[k] = reinterpret_cast<uint4 *>(dst + me * pc * h->rb);
                    a.rec_n16[k] = static_cast<uint32_t>(static_cast<size_t>(sl.h_counts[p]) * h->rb / 16);
                    a.slot_src[k] = sl.vslots + send_off[p]; a.slot_dst[k] = reinterpret_cast<uint32_t *>(dst + l.slots_off + me * pc * 4); a.slot_n[k] = sl.h_counts[p];
                    a.bin_src[k] = sl.bins + static_cast<size_t>(p) * h->bps; a.bin_dst[k] = reinterpret_cast<uint32_t *>(dst + l.bins_off + me * bin_bytes);
                    k++;
                }
                a.bin_n = h->bps;
                k_mg_push<<<static_cast<uint32_t>(k) * MG_PUSH_CTAS, MG_PUSH_THREADS, 0, h->cs>>>(a);
                CK(cudaGetLastError());
            } else { // several streams (copy engines, different NVLink destinations), forked from the communication stream and joined back
                CK(cudaEventRecord(sl.ev_fork, h->cs));
                const int nst = std::min(wfb_mg::CE_STREAMS, n - 1);
                for (int i = 0; i < nst; i++) CK(cudaStreamWaitEvent(h->ce_s[i], sl.ev_fork, 0));
                int k = 0;
                for (int p = 0; p < n; p++) {
                    if (p == h->rank) continue;
                    const CeLayout l = ce_layout(h->peer_cap[p], n, h->rb, h->bps);
                    unsigned char *dst = h->slot[slot_idx].peer[p];
                    const uint64_t me = static_cast<uint64_t>(h->rank), pc = h->peer_cap[p];
                    cudaStream_t st = h->ce_s[k++ % nst];
                    CK(cudaMemcpyAsync(dst + me * pc * h->rb, sl.regions + send_off[p] * h->rb, static_cast<size_t>(sl.h_counts[p]) * h->rb, cudaMemcpyDeviceToDevice, st));
                    CK(cudaMemcpyAsync(dst + l.slots_off + me * pc * 4, sl.vslots + send_off[p], static_cast<size_t>(sl.h_counts[p]) * 4, cudaMemcpyDeviceToDevice, st));
                    CK(cudaMemcpyAsync(dst + l.bins_off + me * bin_bytes, sl.bins + static_cast<size_t>(p) * h->bps, bin_bytes, cudaMemcpyDeviceToDevice, st));
                }
                for (int i = 0; i < nst; i++) { CK(cudaEventRecord(h->ce_ev[i], h->ce_s[i])); CK(cudaStreamWaitEvent(h->cs, h->ce_ev[i], 0)); }
            }
            if (h->trace) CK(cudaEventRecord(sl.tr[5], h->cs));
            // completion tokens: a peer's token arrives after its copies (its stream order) and after the window update it issued last
            // (so that what this rank pushes NEXT into that peer's buffers overwrites nothing still being read)
            if (h->last_done) CK(cudaStreamWaitEvent(h->cs, h->last_done, 0));
            NK(nccl().GroupStart());
            for (int p = 0; p < n; p++) {
                if (p == h->rank) continue;
                NK(nccl().Send(h->tok, 4, NCCL_UINT8, p, h->comm, h->cs));
                NK(nccl().Recv(h->tok + 1 + p, 4, NCCL_UINT8, p, h->comm, h->cs));
            }
            NK(nccl().GroupEnd());
            CK(cudaEventRecord(sl.ev_a2a, h->cs));
            return 0;
        }
        const size_t need = std::max<size_t>(1, tot);
        if (sl.recv_bytes < need * h->rb || sl.recv_slots_cap < need) {
            CK(cudaDeviceSynchronize());
            cudaFree(sl.recv); cudaFree(sl.recv_slots);
            sl.recv_slots_cap = need * 5 / 4; sl.recv_bytes = sl.recv_slots_cap * h->rb;
            CK(cudaMalloc(&sl.recv, sl.recv_bytes));
            CK(cudaMalloc(&sl.recv_slots, sizeof(uint32_t) * sl.recv_slots_cap));
        }
        if (sl.done_recorded) CK(cudaStreamWaitEvent(h->cs, sl.ev_done, 0)); // the window update that read these receive buffers two steps ago
        if (h->trace) CK(cudaEventRecord(sl.tr[4], h->cs));
        size_t send_off[MAX_SHARDS + 1]; send_off[0] = 0;
        for (int p = 0; p < n; p++) send_off[p + 1] = send_off[p] + sl.h_counts[p];
        const size_t bin_bytes = sizeof(uint32_t) * h->bps;
        if (n > 1) {
            {   // this rank's own share does not go through NCCL: plain copies on the copy engine, next to the group
                const int p = h->rank;
                CK(cudaStreamWaitEvent(h->cs2, sl.ev_src, 0));
                if (sl.done_recorded) CK(cudaStreamWaitEvent(h->cs2, sl.ev_done, 0));
                CK(cudaMemcpyAsync(sl.recv + static_cast<size_t>(sl.offs[p]) * h->rb, sl.regions + send_off[p] * h->rb, static_cast<size_t>(sl.h_counts[p]) * h->rb, cudaMemcpyDeviceToDevice, h->cs2));
                CK(cudaMemcpyAsync(sl.recv_slots + sl.offs[p], sl.vslots + send_off[p], static_cast<size_t>(sl.h_counts[p]) * 4, cudaMemcpyDeviceToDevice, h->cs2));
                CK(cudaMemcpyAsync(sl.recv_bins + static_cast<size_t>(p) * h->bps, sl.bins + static_cast<size_t>(p) * h->bps, bin_bytes, cudaMemcpyDeviceToDevice, h->cs2));
                CK(cudaEventRecord(sl.ev_self, h->cs2));
            }
            NK(nccl().GroupStart());
            for (int p = 0; p < n; p++) {
                if (p == h->rank) continue;
                NK(nccl().Send(sl.regions + send_off[p] * h->rb, static_cast<size_t>(sl.h_counts[p]) * h->rb, NCCL_UINT8, p, h->comm, h->cs));
                NK(nccl().Recv(sl.recv + static_cast<size_t>(sl.offs[p]) * h->rb, static_cast<size_t>(sl.h_recv[2 * p]) * h->rb, NCCL_UINT8, p, h->comm, h->cs));
                NK(nccl().Send(sl.vslots + send_off[p], static_cast<size_t>(sl.h_counts[p]) * 4, NCCL_UINT8, p, h->comm, h->cs));
                NK(nccl().Recv(sl.recv_slots + sl.offs[p], static_cast<size_t>(sl.h_recv[2 * p]) * 4, NCCL_UINT8, p, h->comm, h->cs));
                NK(nccl().Send(sl.bins + static_cast<size_t>(p) * h->bps, bin_bytes, NCCL_UINT8, p, h->comm, h->cs));
                NK(nccl().Recv(sl.recv_bins + static_cast<size_t>(p) * h->bps, bin_bytes, NCCL_UINT8, p, h->comm, h->cs));
            }
            NK(nccl().GroupEnd());
        } else {
            CK(cudaMemcpyAsync(sl.recv, sl.regions, static_cast<size_t>(sl.h_counts[0]) * h->rb, cudaMemcpyDeviceToDevice, h->cs));
            CK(cudaMemcpyAsync(sl.recv_slots, sl.vslots, static_cast<size_t>(sl.h_counts[0]) * 4, cudaMemcpyDeviceToDevice, h->cs));
            CK(cudaMemcpyAsync(sl.recv_bins, sl.bins, bin_bytes, cudaMemcpyDeviceToDevice, h->cs));
        }
        CK(cudaEventRecord(sl.ev_a2a, h->cs));
        if (h->trace) CK(cudaEventRecord(sl.tr[5], h->cs));
        return 0;
    }
    size_t tiles = 0; // every source's chunk at its tile position of the receive buffer (read in place)
    for (int p = 0; p < n; p++) { sl.offs[p] = static_cast<uint32_t>(tiles * TILE); tiles += (static_cast<size_t>(sl.h_recv[2 * p]) + TILE - 1) / TILE; sl.wms[p] = sl.h_recv[2 * p + 1]; }
    if (tiles * TILE > 0x7fffffffull) return WFB_E_CAPACITY;
    const size_t need = std::max<size_t>(1, tiles * TILE) * h->rb;
    if (sl.recv_bytes < need) {
        CK(cudaDeviceSynchronize());
        cudaFree(sl.recv);
        sl.recv_bytes = need * 5 / 4;
        CK(cudaMalloc(&sl.recv, sl.recv_bytes));
    }
    if (sl.done_recorded) CK(cudaStreamWaitEvent(h->cs, sl.ev_done, 0)); // the window update that read this receive buffer two steps ago
    if (h->trace) CK(cudaEventRecord(sl.tr[4], h->cs));
    if (n > 1) {
        NK(nccl().GroupStart());
        for (int p = 0; p < n; p++) {
            NK(nccl().Send(sl.regions + static_cast<size_t>(p) * sl.region_cap * h->rb, static_cast<size_t>(sl.h_counts[p]) * h->rb, NCCL_UINT8, p, h->comm, h->cs));
            NK(nccl().Recv(sl.recv + static_cast<size_t>(sl.offs[p]) * h->rb, static_cast<size_t>(sl.h_recv[2 * p]) * h->rb, NCCL_UINT8, p, h->comm, h->cs));
        }
        NK(nccl().GroupEnd());
    } else CK(cudaMemcpyAsync(sl.recv, sl.regions, static_cast<size_t>(sl.h_counts[0]) * h->rb, cudaMemcpyDeviceToDevice, h->cs));
    CK(cudaEventRecord(sl.ev_a2a, h->cs));
    if (h->trace) CK(cudaEventRecord(sl.tr[5], h->cs));
    return 0;
}

// window update on what mg_exchange delivered (caller's stream)
static int mg_update(wfb_mg *h, MgSlot &sl, void *out, uint64_t *out_ts, uint32_t out_cap, uint32_t *n_out_dev, cudaStream_t s, bool append = false)
{
    const int n = h->nranks;
    CK(cudaStreamWaitEvent(s, sl.ev_a2a, 0));
    if (h->bucketed && n > 1) CK(cudaStreamWaitEvent(s, sl.ev_self, 0));
    if (h->trace) CK(cudaEventRecord(sl.tr[2], s));
    uint64_t items = 0;
    for (int p = 0; p < n; p++) items += sl.h_recv[2 * p];
    if (append && items != 0) { k_mg_pre_append<<<1, 1, 0, s>>>(h->ffat->ff.results_total, n_out_dev); CK(cudaGetLastError()); }
    int rc;
    if (h->bucketed) {
        rc = ffat_process_prebucketed(h->ffat, sl.recv, sl.recv_slots, sl.recv_bins, static_cast<uint32_t>(n), h->bps, sl.offs, sl.wms, h->shard_slots - 1u, h->shift,
                                      out, out_ts, out_cap, n_out_dev, s, append, static_cast<uint32_t>(items));
    } else {
        h->chunks.resize(n);
        for (int p = 0; p < n; p++) { // source-rank order = global stream order
            wfb_batch_t &b = h->chunks[p];
            b.tuples = sl.recv + static_cast<size_t>(sl.offs[p]) * h->rb; b.ts = nullptr; b.watermark = sl.wms[p]; b.n = static_cast<uint32_t>(sl.h_recv[2 * p]); b.reserved = 0;
        }
        h->ffat->append_results = append;
        rc = wfb_ffat_process_cb(h->ffat, nullptr, h->chunks.data(), static_cast<uint32_t>(n), out, out_ts, out_cap, n_out_dev, s);
        h->ffat->append_results = false;
    }
    if (rc) return rc;
    CK(cudaEventRecord(sl.ev_done, s));
    h->last_done = sl.ev_done;
    if (h->trace) { CK(cudaEventRecord(sl.tr[3], s)); sl.tr_valid = true; }
    sl.done_recorded = true;
    return 0;
}

int wfb_mg_step(wfb_mg_t *h, const wfb_functors_t *pre, const wfb_batch_t *batches_h, uint32_t nbatches, uint64_t watermark,
                void *out_results, uint64_t *out_ts, uint32_t out_capacity, uint32_t *n_out_dev, void *stream)
{
    if (!h || !n_out_dev || (nbatches && !batches_h)) return WFB_E_BADARG;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    MgSlot &cur = h->slot[h->step_no % 4];
    h->step_no++;
    // call i: the records of step i-2 start travelling (communication stream; their sizes reached the host a step ago), the source pass
    // of step i runs next to them, then the window update of step i-3 -- whose records arrived during the previous call, so the compute
    // stream never waits for an exchange, and no rank waits for the slowest peer of the step in flight
    MgSlot *ex = h->npend >= 2 ? h->pend[h->npend - 2] : nullptr, *upd = h->npend == 3 ? h->pend[0] : nullptr;
    int rc;
    const double t0 = h->trace ? host_now_us() : 0;
    if (ex != nullptr && !ex->exchanged) { rc = mg_exchange(h, *ex); if (rc) return rc; ex->exchanged = true; }
    const double t1 = h->trace ? host_now_us() : 0;
    rc = mg_source(h, cur, pre, batches_h, nbatches, watermark, s); if (rc) return rc;
    const double t2 = h->trace ? host_now_us() : 0;
    cur.exchanged = false;
    if (upd != nullptr) { h->pend[0] = h->pend[1]; h->pend[1] = h->pend[2]; h->pend[2] = &cur; } else h->pend[h->npend++] = &cur;
    if (upd == nullptr) { CK(cudaMemsetAsync(n_out_dev, 0, sizeof(uint32_t), s)); return 0; }
    rc = mg_update(h, *upd, out_results, out_ts, out_capacity, n_out_dev, s);
    upd->exchanged = false;
    if (h->trace) { // host time spent ISSUING the three parts of a step (includes any wait for a pinned staging slot)
        const double t3 = host_now_us();
        h->host_acc[0] += t1 - t0; h->host_acc[1] += t2 - t1; h->host_acc[2] += t3 - t2;
        if (++h->host_n % 64 == 0) {
            std::fprintf(stderr, "[wfb_mg rank %d] host us/step over 64 steps: issue exchange %.0f | issue source + sizes %.0f | issue update %.0f\n", h->rank,
                         h->host_acc[0] / 64, h->host_acc[1] / 64, h->host_acc[2] / 64);
            h->host_acc[0] = h->host_acc[1] = h->host_acc[2] = 0;
        }
    }
    return rc;
}

int wfb_mg_flush(wfb_mg_t *h, void *out_results, uint64_t *out_ts, uint32_t out_capacity, uint32_t *n_out_dev, void *stream)
{
    if (!h || !n_out_dev) return WFB_E_BADARG;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (h->npend == 0) { CK(cudaMemsetAsync(n_out_dev, 0, sizeof(uint32_t), s)); return 0; }
    const int n = h->npend; h->npend = 0;
    int rc;
    for (int i = 0; i < n; i++) { // oldest first; the results of the later steps follow the first one's in the buffer
        MgSlot &sl = *h->pend[i];
        if (!sl.exchanged) { rc = mg_exchange(h, sl); if (rc) return rc; }
        rc = mg_update(h, sl, out_results, out_ts, out_capacity, n_out_dev, s, i != 0); if (rc) return rc;
        sl.exchanged = false;
        h->pend[i] = nullptr;
    }
    return 0;
}

uint64_t wfb_mg_launches(const wfb_mg_t *h) { return h ? wfb_engine_launches(h->eng) + wfb_ffat_launches(h->ffat) : 0; }

int wfb_mg_stats(wfb_mg_t *h, uint32_t *err_flags_h, uint64_t *results_total_h, void *stream)
{
    if (!h) return WFB_E_BADARG;
    uint32_t nk = 0, ef = 0; uint64_t tot = 0;
    int rc = wfb_ffat_stats(h->ffat, &nk, &ef, stream); if (rc) return rc;
    rc = wfb_ffat_results_total(h->ffat, &tot, stream); if (rc) return rc;
    if (err_flags_h) *err_flags_h = ef;
    if (results_total_h) *results_total_h = tot;
    return 0;
}

int wfb_gen_tuple64(uint64_t seed, uint64_t start, uint32_t n, int key_mode, uint64_t nkeys, const double *zipf_cdf,
                    void *tuples, uint64_t *ts, void *stream)
{
    int rc = device_ready(); if (rc) return rc;
    if ((n && !tuples) || nkeys == 0 || (key_mode == 2 && !zipf_cdf)) return WFB_E_BADARG;
    if (n == 0) return 0;
    const uint32_t grid = std::min((n + 255) / 256, static_cast<uint32_t>(g_num_sms) * 16u);
    k_gen_tuple64<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(seed, start, n, key_mode, nkeys, zipf_cdf,
                                                                       static_cast<wfb_tuple64_t *>(tuples), ts);
    CK(cudaGetLastError());
    return 0;
}

} // extern "C"

#ifdef WFB_BK_TRACE
// debug build only: phase timestamps (globaltimer ns) of the last k_ffat_update_buckets launch, 8 per CTA
extern "C" int wfb_debug_bk_trace(unsigned long long *out) { return cudaMemcpyFromSymbol(out, wfb::g_bk_trace, sizeof(unsigned long long) * 1024 * 8) == cudaSuccess ? 0 : -1; }
#endif
