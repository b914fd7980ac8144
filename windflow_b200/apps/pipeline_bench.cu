// pipeline_bench.cu -- the headline pipeline of BASELINE.json written against the builder API (include/wf/windflow_gpu.hpp):
//
//     SourceGPU (a ring of batches resident in HBM) -> Map_GPU -> Filter_GPU -> Ffat_Windows_GPU (count-based) -> Sink
//
// The same workload as bench.py (64-byte tuples of the seeded synthetic stream, batch 65536, 65536 uniform keys, map ivalue += 2 /
// fvalue *= 1.0000001, filter (ivalue & 1) == 0, windows 4096 / 64, Nb 65) -- but driven the way an application drives it: operators
// built with the builders, wired with MultiPipe, every replica a thread of the runtime with a queue in front of it. The window
// replica takes up to K queued batches per svc() (withMaxBatchesPerCall): K = 1 is the reference's one-batch-per-svc behaviour,
// larger K is what a replica finds queued when the source is faster than one launch sequence.
//   usage: pipeline_bench.bin [K=128] [timed_batches=40960] [keys=65536] [nb=65] [ring_batches=512] [sink_replicas=2] [style=fluent]
// style: "fluent" = source.chain(map).chain(filter).add(ffat) in ONE expression -- the functor types reach the window operator's
// program and are inlined into its tile pass; "statements" = one mp.chain(...) per statement -- the same fusion through per-stage
// device function pointers (the types are gone by the time the window operator arrives).
// prints one JSON line. Build: nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr
//        --expt-extended-lambda -I include windflow_b200/apps/pipeline_bench.cu -L windflow_b200 -lwfb200
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <optional>
#include <thread>
#include <wf/windflow_gpu.hpp>

using namespace wf;

struct tuple64_t { uint64_t key, id; int64_t ivalue; double fvalue; uint64_t pad[4]; };
struct result32_t {
    uint64_t key, id; int64_t isum; double fsum;
    __host__ __device__ result32_t(): key(0), id(0), isum(0), fsum(0) {}
    __host__ __device__ result32_t(uint64_t k, uint64_t i): key(k), id(i), isum(0), fsum(0) {}
};
struct MapF { __host__ __device__ void operator()(tuple64_t &t) { t.ivalue += 2; t.fvalue *= 1.0000001; } };
struct FiltF { __host__ __device__ bool operator()(tuple64_t &t) { return (t.ivalue & 1) == 0; } };
struct LiftF { __host__ __device__ void operator()(const tuple64_t &t, result32_t &r) { r.key = t.key; r.id = 0; r.isum = t.ivalue; r.fsum = t.fvalue; } };
struct CombF { __host__ __device__ void operator()(const result32_t &a, const result32_t &b, result32_t &o) { o.isum = a.isum + b.isum; o.fsum = a.fsum + b.fsum; } };
struct KeyF { __host__ __device__ uint64_t operator()(const tuple64_t &t) { return t.key; } };

static std::atomic<uint64_t> g_seen_wm{0};      // watermark (= index of the last input batch) of the newest result batch the sink has seen
static std::atomic<uint64_t> g_windows{0};
static std::atomic<long long> g_isum{0};
// every sink replica keeps its own totals and adds them to the global ones at end of stream (the reference's sinks are per-tuple functors too)
struct SinkF {
    uint64_t windows = 0; long long isum = 0;
    void operator()(std::optional<result32_t> &r)
    {
        if (r) { windows++; isum += r->isum; }
        else { g_windows.fetch_add(windows); g_isum.fetch_add(isum); }
    }
};

int main(int argc, char **argv)
{
    const size_t K = argc > 1 ? std::strtoull(argv[1], nullptr, 10) : 128;
    const uint64_t timed = argc > 2 ? std::strtoull(argv[2], nullptr, 10) : 40960;
    const uint64_t nkeys = argc > 3 ? std::strtoull(argv[3], nullptr, 10) : 65536;
    const size_t nb = argc > 4 ? std::strtoull(argv[4], nullptr, 10) : 65;
    const uint64_t ring = argc > 5 ? std::strtoull(argv[5], nullptr, 10) : 512; // 512 x 4 MiB = 2 GiB of input, far larger than L2
    const size_t sinks = argc > 6 ? std::strtoull(argv[6], nullptr, 10) : 2;
    const bool fluent = !(argc > 7 && std::strcmp(argv[7], "statements") == 0);
    const uint64_t BATCH = 65536, WIN = 4096, SLIDE = 64;
    const uint64_t B = (nb - 1) * SLIDE + WIN;
    // every key past its first trigger before the clock starts: B surviving tuples per key at selectivity 0.5
    const uint64_t prime = (2 * B * nkeys + BATCH - 1) / BATCH + 2 * ring;

    // the ring: `ring` consecutive batches of the synthetic stream, generated once on the device
    tuple64_t *d_tuples = nullptr; uint64_t *d_ts = nullptr;
    gpuErrChk(cudaMalloc(&d_tuples, ring * BATCH * sizeof(tuple64_t)));
    gpuErrChk(cudaMalloc(&d_ts, ring * BATCH * sizeof(uint64_t)));
    for (uint64_t r = 0; r < ring; r++)
        wfbErrChk(wfb_gen_tuple64(0x5EED5EEDull, r * BATCH, static_cast<uint32_t>(BATCH), 1, nkeys, nullptr, d_tuples + r * BATCH, d_ts + r * BATCH, nullptr));
    gpuErrChk(cudaDeviceSynchronize());

    std::chrono::steady_clock::time_point t0, t1;
    auto source = [&](SourceGPU_Shipper<tuple64_t> &sh) {
        auto push = [&](uint64_t i) { const uint64_t r = i % ring; sh.pushBatch(d_tuples + r * BATCH, d_ts + r * BATCH, BATCH, i + 1); };
        auto drained = [&](uint64_t upto) { while (g_seen_wm.load(std::memory_order_acquire) < upto) std::this_thread::sleep_for(std::chrono::microseconds(50)); };
        for (uint64_t i = 0; i < prime; i++) push(i);
        drained(prime);
        t0 = std::chrono::steady_clock::now();
        for (uint64_t i = prime; i < prime + timed; i++) push(i);
        drained(prime + timed);
        t1 = std::chrono::steady_clock::now();
    };

    size_t threads = 0;
    {
        PipeGraph graph("pipeline_bench", Execution_Mode_t::DEFAULT, Time_Policy_t::EVENT_TIME);
        MultiPipe &mp = graph.add_source(SourceGPU_Builder(source).withName("source").build());
        auto ffat = Ffat_WindowsGPU_Builder(LiftF(), CombF()).withName("ffat").withKeyBy(KeyF()).withCBWindows(WIN, SLIDE).withNumWinPerBatch(nb)
                        .withMaxKeys(static_cast<uint32_t>(nkeys)).withDenseKeys().withMaxBatchesPerCall(K).build();
        if (fluent) mp.chain(MapGPU_Builder(MapF()).withName("map").build()).chain(FilterGPU_Builder(FiltF()).withName("filter").build()).add(ffat);
        else {
            mp.chain(MapGPU_Builder(MapF()).withName("map").build());
            mp.chain(FilterGPU_Builder(FiltF()).withName("filter").build());
            mp.add(ffat);
        }
        // the sink also publishes how far the stream has been processed (the watermark of every result batch = its last input batch)
        mp.chain_sink(Sink_Builder(SinkF()).withName("sink").withParallelism(sinks).withWatermarkProbe(&g_seen_wm).build());
        threads = graph.getNumThreads();
        graph.run();
    }
    const double sec = std::chrono::duration<double>(t1 - t0).count();
    std::printf("{\"api\": \"facade\", \"max_batches_per_call\": %zu, \"tuples\": %llu, \"seconds\": %.6f, \"tuples_per_s\": %.1f, \"windows\": %llu, \"isum\": %lld, "
                "\"threads\": %zu, \"keys\": %llu, \"nb\": %zu, \"primed_batches\": %llu, \"style\": \"%s\"}\n",
                K, static_cast<unsigned long long>(timed * BATCH), sec, timed * BATCH / sec, static_cast<unsigned long long>(g_windows.load()),
                g_isum.load(), threads, static_cast<unsigned long long>(nkeys), nb, static_cast<unsigned long long>(prime), fluent ? "fluent" : "statements");
    cudaFree(d_tuples); cudaFree(d_ts);
    return 0;
}
