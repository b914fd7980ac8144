// Application-level test of the WindFlow builder API over libwfb200 (include/wf/windflow_gpu.hpp), written like the
// reference's own GPU tests (tests/graph_tests_gpu/test_graph_gpu_1.cpp, tests/win_tests_gpu/test_win_fat_gpu_tb.cpp):
// same tuple / result structs and functors (graph_common_gpu.hpp, win_common_gpu.hpp), a Sink accumulating a global sum,
// and -- since the stream is deterministic (value = i per key) -- a closed-form expected value instead of run-to-run
// invariance. Prints FACADE_OK on success.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <optional>
#include <wf/windflow_gpu.hpp>

using namespace wf;

struct tuple_t { // tests/win_tests_gpu/win_common_gpu.hpp:40-58
    size_t key; uint64_t id; int64_t value;
    __host__ __device__ tuple_t(): key(0), id(0), value(0) {}
    __host__ __device__ tuple_t(size_t k, uint64_t i): key(k), id(i), value(0) {}
};
struct result_t { // :61-80
    size_t key; uint64_t id; int64_t value;
    __host__ __device__ result_t(): key(0), id(0), value(0) {}
    __host__ __device__ result_t(size_t k, uint64_t i): key(k), id(i), value(0) {}
};

struct Source_Positive_Functor { // :100-116 (timestamps increase by 1 here instead of a random gap)
    size_t len, keys;
    void operator()(Source_Shipper<tuple_t> &shipper)
    {
        uint64_t next_ts = 0;
        for (size_t i = 1; i <= len; i++)
            for (size_t k = 0; k < keys; k++) {
                tuple_t t(k, 0); t.value = static_cast<int64_t>(i);
                shipper.pushWithTimestamp(t, next_ts); shipper.setNextWatermark(next_ts); next_ts++;
            }
    }
};
struct Map_Functor_GPU { __host__ __device__ void operator()(tuple_t &t) { t.value = t.value + 2; } };              // :221-229
struct Filter_Functor_GPU { int mod; __host__ __device__ bool operator()(tuple_t &t) { return t.value % mod == 0; } }; // graph_common_gpu.hpp:198-215
struct Lift_Functor_GPU { __host__ __device__ void operator()(const tuple_t &t, result_t &r) { r.value = t.value; } }; // :295-303
struct Comb_Functor_GPU { __host__ __device__ void operator()(const result_t &a, const result_t &b, result_t &o) { o.value = a.value + b.value; } }; // :306-314
struct Key_Functor { __host__ __device__ size_t operator()(const tuple_t &t) { return t.key; } };
struct Reduce_Functor_GPU { // graph_common_gpu.hpp:268-279
    __host__ __device__ tuple_t operator()(const tuple_t &a, const tuple_t &b) { tuple_t r; r.key = a.key; r.value = a.value + b.value; return r; }
};

struct map_state_t { int64_t counter; __host__ __device__ map_state_t(): counter(0) {} };       // graph_common_gpu.hpp:52-60
struct filter_state_t { int64_t counter; __host__ __device__ filter_state_t(): counter(0) {} }; // :63-71
struct Map_Functor_GPU_KB { __host__ __device__ void operator()(tuple_t &t, map_state_t &state) { state.counter++; t.value += state.counter; } }; // :256-265
struct Filter_Functor_GPU_KB { // :221-231 (with a predicate on the updated value so that the compaction is exercised)
    __host__ __device__ bool operator()(tuple_t &t, filter_state_t &state) { state.counter++; t.value += state.counter; return (t.value & 1) == 0; }
};

static std::atomic<long> global_sum{0};
static std::atomic<long> received{0};
struct Sink_Functor {
    void operator()(std::optional<result_t> &out) { if (out) { global_sum += out->value; received++; } }
};
struct Sink_Functor_T {
    void operator()(std::optional<tuple_t> &out) { if (out) { global_sum += out->value; received++; } }
};

static long win_count[16], win_sum[16]; static bool win_order_ok = true;
struct Sink_Functor_Win { // windows of a key must arrive with consecutive ids 0, 1, 2, ...
    void operator()(std::optional<result_t> &out)
    {
        if (!out) return;
        if (static_cast<long>(out->id) != win_count[out->key]) win_order_ok = false;
        win_count[out->key]++; win_sum[out->key] += out->value; received++;
    }
};

static void check(const char *what, long got, long exp)
{
    if (got != exp) { std::printf("FAILED %s: got %ld expected %ld\n", what, got, exp); std::exit(1); }
    std::printf("%s OK (%ld)\n", what, got);
}

// compile-time only (never called): the ways an application may hold on to what chain() returns keep compiling now that the
// chain() of a stateless operator returns a proxy (FusedPipe) instead of MultiPipe &
[[maybe_unused]] static void api_shapes(PipeGraph &graph, Source_Positive_Functor sf)
{
    MultiPipe &mp = graph.add_source(Source_Builder(sf).withName("source").withOutputBatchSize(64).build());
    MultiPipe &same = mp.chain(MapGPU_Builder(Map_Functor_GPU()).withName("m").build());                      // reference bound to the proxy's MultiPipe
    same.chain(FilterGPU_Builder(Filter_Functor_GPU{2}).withName("f").build()).chain_sink(Sink_Builder(Sink_Functor_T()).withName("s").build()); // fluent into a sink
    MultiPipe &mp2 = graph.add_source(Source_Builder(sf).withName("source2").withOutputBatchSize(64).build());
    mp2.add(MapGPU_Builder(Map_Functor_GPU()).withName("m2").build())
       .add(ReduceGPU_Builder(Reduce_Functor_GPU()).withName("r2").withKeyBy(Key_Functor()).build())                // stateless run, then an operator that is not fused
       .add_sink(Sink_Builder(Sink_Functor_T()).withName("s2").build());
    (void) mp2.getNumThreads();
}

int main()
{
    const size_t len = 3000, keys = 7, batch = 1000;
    // ---- test 1: Source -> Map_GPU -> Filter_GPU -> Sink (test_graph_gpu_1 shape) -------------------------------------
    {
        global_sum = 0; received = 0;
        PipeGraph graph("test_graph_gpu", Execution_Mode_t::DEFAULT, Time_Policy_t::EVENT_TIME);
        Source_Positive_Functor sf{len, keys};
        MultiPipe &mp = graph.add_source(Source_Builder(sf).withName("source").withParallelism(1).withOutputBatchSize(batch).build());
        mp.chain(MapGPU_Builder(Map_Functor_GPU()).withName("mapgpu1").withParallelism(2).build());
        mp.chain(FilterGPU_Builder(Filter_Functor_GPU{3}).withName("filtergpu1").build());
        mp.chain(MapGPU_Builder(Map_Functor_GPU()).withName("mapgpu2").build());
        mp.chain_sink(Sink_Builder(Sink_Functor_T()).withName("sink").build());
        graph.run();
        long exp = 0, cnt = 0;
        for (size_t i = 1; i <= len; i++) if ((i + 2) % 3 == 0) { exp += static_cast<long>(i + 4) * keys; cnt += keys; }
        check("map-filter-map sum", global_sum, exp); check("map-filter-map count", received, cnt);
    }
    // ---- test 2: Source -> Map_GPU -> Ffat_Windows_GPU (CB, keyed) -> Sink ------------------------------------------------
    {
        global_sum = 0; received = 0;
        const uint64_t win = 64, slide = 16; const size_t nwb = 3;
        PipeGraph graph("test_win_fat_gpu_cb", Execution_Mode_t::DEFAULT, Time_Policy_t::EVENT_TIME);
        Source_Positive_Functor sf{len, keys};
        MultiPipe &mp = graph.add_source(Source_Builder(sf).withName("source").withOutputBatchSize(batch).build());
        mp.chain(MapGPU_Builder(Map_Functor_GPU()).withName("mapgpu").build());
        mp.add(Ffat_WindowsGPU_Builder(Lift_Functor_GPU(), Comb_Functor_GPU()).withName("ffat_agg").withKeyBy(Key_Functor())
                   .withCBWindows(win, slide).withNumWinPerBatch(nwb).withMaxKeys(16).build());
        mp.chain_sink(Sink_Builder(Sink_Functor()).withName("sink").build());
        graph.run();
        // per key: items value = i + 2; groups fired = 1 + (len - B) / (slide*nwb), B = (nwb-1)*slide + win
        const uint64_t B = (nwb - 1) * slide + win;
        const uint64_t groups = 1 + (len - B) / (slide * nwb);
        long exp = 0;
        for (uint64_t g = 0; g < groups * nwb; g++) { const long a = g * slide + 1 + 2, b = g * slide + win + 2; exp += (a + b) * static_cast<long>(win) / 2; }
        check("ffat cb windows sum", global_sum, exp * static_cast<long>(keys));
        check("ffat cb windows count", received, static_cast<long>(groups * nwb * keys));
    }
    // ---- test 2b: the same pipeline written as ONE expression: the Map/Filter functor types become part of the window operator's
    // program (FusedPipe / TypedChain: no thunks); a filter that keeps everything rides along. Same expected sums.
    {
        global_sum = 0; received = 0;
        const uint64_t win = 64, slide = 16; const size_t nwb = 3;
        PipeGraph graph("test_win_fat_gpu_cb_fluent", Execution_Mode_t::DEFAULT, Time_Policy_t::EVENT_TIME);
        Source_Positive_Functor sf{len, keys};
        graph.add_source(Source_Builder(sf).withName("source").withOutputBatchSize(batch).build())
            .chain(MapGPU_Builder(Map_Functor_GPU()).withName("mapgpu").build())
            .chain(FilterGPU_Builder(Filter_Functor_GPU{1}).withName("filtergpu_all").build())
            .add(Ffat_WindowsGPU_Builder(Lift_Functor_GPU(), Comb_Functor_GPU()).withName("ffat_agg").withKeyBy(Key_Functor())
                     .withCBWindows(win, slide).withNumWinPerBatch(nwb).withMaxKeys(16).build())
            .chain_sink(Sink_Builder(Sink_Functor()).withName("sink").build());
        graph.run();
        const uint64_t B = (nwb - 1) * slide + win;
        const uint64_t groups = 1 + (len - B) / (slide * nwb);
        long exp = 0;
        for (uint64_t g = 0; g < groups * nwb; g++) { const long a = g * slide + 1 + 2, b = g * slide + win + 2; exp += (a + b) * static_cast<long>(win) / 2; }
        check("ffat cb windows sum (fluent, typed fusion)", global_sum, exp * static_cast<long>(keys));
        check("ffat cb windows count (fluent, typed fusion)", received, static_cast<long>(groups * nwb * keys));
    }
    // ---- test 3: Source -> Reduce_GPU (keyed) -> Sink (test_graph_gpu_4 shape) -----------------------------------------------
    {
        global_sum = 0; received = 0;
        PipeGraph graph("test_reduce_gpu", Execution_Mode_t::DEFAULT, Time_Policy_t::EVENT_TIME);
        Source_Positive_Functor sf{len, keys};
        MultiPipe &mp = graph.add_source(Source_Builder(sf).withName("source").withOutputBatchSize(batch).build());
        mp.chain(ReduceGPU_Builder(Reduce_Functor_GPU()).withName("reducegpu").withKeyBy(Key_Functor()).build());
        mp.chain_sink(Sink_Builder(Sink_Functor_T()).withName("sink").build());
        graph.run();
        long exp = 0;
        for (size_t i = 1; i <= len; i++) exp += static_cast<long>(i) * keys; // a reduce only regroups the values
        check("reduce_by_key sum", global_sum, exp);
        const long nb = (len * keys + batch - 1) / batch;
        check("reduce_by_key items", received, nb * static_cast<long>(keys)); // every batch holds all 7 keys
    }
    // ---- test 4: Source -> Map_GPU (keyed-stateful) -> Filter_GPU (keyed-stateful) -> Sink (test_graph_gpu / merge kb shapes) --
    {
        global_sum = 0; received = 0;
        PipeGraph graph("test_stateful_gpu", Execution_Mode_t::DEFAULT, Time_Policy_t::EVENT_TIME);
        Source_Positive_Functor sf{len, keys};
        MultiPipe &mp = graph.add_source(Source_Builder(sf).withName("source").withOutputBatchSize(batch).build());
        mp.chain(MapGPU_Builder(Map_Functor_GPU_KB()).withName("mapgpu_kb").withKeyBy(Key_Functor()).withMaxKeys(64).build());
        mp.chain(FilterGPU_Builder(Filter_Functor_GPU_KB()).withName("filtergpu_kb").withKeyBy(Key_Functor()).withMaxKeys(64).build());
        mp.chain_sink(Sink_Builder(Sink_Functor_T()).withName("sink").build());
        graph.run();
        // per key: the i-th tuple (value i) gets +i from the map's counter (2i) and +i from the filter's counter (3i); kept when even
        long exp_sum = 0, exp_cnt = 0;
        for (size_t i = 1; i <= len; i++) { const long v = 3 * static_cast<long>(i); if ((v & 1) == 0) { exp_sum += v * keys; exp_cnt += keys; } }
        check("stateful map -> stateful filter sum", global_sum, exp_sum);
        check("stateful map -> stateful filter items", received, exp_cnt);
    }
    // ---- test 5: Source -> Ffat_Windows_GPU keyed, TIME-based windows -> Sink (test_win_fat_gpu_tb shape) ------------------------
    {
        received = 0; win_order_ok = true;
        for (int k = 0; k < 16; k++) { win_count[k] = 0; win_sum[k] = 0; }
        const uint64_t win = 210, slide = 70; // microseconds; timestamps grow by 1 per tuple, so a key sees one tuple every 7 us
        PipeGraph graph("test_win_fat_gpu_tb", Execution_Mode_t::DEFAULT, Time_Policy_t::EVENT_TIME);
        Source_Positive_Functor sf{len, keys};
        MultiPipe &mp = graph.add_source(Source_Builder(sf).withName("source").withOutputBatchSize(batch).build());
        mp.add(Ffat_WindowsGPU_Builder(Lift_Functor_GPU(), Comb_Functor_GPU()).withName("ffat_tb").withKeyBy(Key_Functor())
                   .withTBWindows(std::chrono::microseconds(win), std::chrono::microseconds(slide)).withNumWinPerBatch(3).withMaxKeys(16).build());
        mp.chain_sink(Sink_Builder(Sink_Functor_Win()).withName("sink").build());
        graph.run();
        if (!win_order_ok) { std::printf("FAILED tb windows: ids not consecutive per key\n"); return 1; }
        long fired = 0;
        for (size_t k = 0; k < keys; k++) { // window g of key k = sum of the values of its tuples with ts in [g*slide, g*slide + win)
            long exp = 0;
            for (long g = 0; g < win_count[k]; g++)
                for (size_t i = 1; i <= len; i++) { const uint64_t ts = (i - 1) * keys + k; if (ts >= g * slide && ts < g * slide + win) exp += static_cast<long>(i); }
            if (exp != win_sum[k]) { std::printf("FAILED tb windows of key %zu: got %ld expected %ld over %ld windows\n", k, win_sum[k], exp, win_count[k]); return 1; }
            if (win_count[k] != win_count[0]) { std::printf("FAILED tb windows: keys fired different numbers of windows\n"); return 1; }
            fired += win_count[k];
        }
        // the stream spans len*keys = 21000 us: all but the last few groups of 3 windows have fired
        if (fired < static_cast<long>(keys) * 270 || fired % 3 != 0) { std::printf("FAILED tb windows: %ld fired\n", fired); return 1; }
        std::printf("ffat tb windows OK (%ld windows, %ld per key)\n", fired, win_count[0]);
    }
    std::printf("FACADE_OK\n");
    return 0;
}
