// oracle/ref_pipeline.cu -- TEST INFRASTRUCTURE: the UNMODIFIED reference (wf/windflow.hpp + wf/windflow_gpu.hpp, included from
// /root/reference at build time; FastFlow's API is provided by include/ff/, TBB's map by oracle/shim/) running the hot path's
// pipelines with the bench schema of SURVEY.md 8d. Nothing of the product links or loads this program; tests and bench.py run
// it as a checker / as the "reference on this box" arm.
//
//   ref_pipeline <mode> key=value ...
//   modes   gpu_cb   Source -> Map_GPU -> Filter_GPU -> Ffat_Windows_GPU (count-based)      [reference GPU operators]
//           gpu_tb   the same with time-based windows (win / slide / lateness in timestamp units)
//           cpu_cb   Source -> Map -> Filter -> Ffat_Windows (count-based), `par` replicas  [reference CPU operators, BASELINE cfg 1]
//           gpu_mf   Source -> Map_GPU -> Filter_GPU -> Sink                               (BASELINE cfg 2)
//           gpu_red  Source -> Reduce_GPU keyed -> Sink                                     (BASELINE cfg 3)
//   keys    in=<file>     stream to replay: u64 n, then n x tuple64, n x u64 timestamps, n x u64 watermarks
//           out=<file>    results: u64 m, then m x {result32, u64 ts} (or tuple64 + ts for gpu_mf / gpu_red), unordered
//           gen=<n>       instead of in=: generate n tuples of the synthetic stream (seed 0x5EED5EED, uniform keys)
//           det=1 (cpu_cb: DETERMINISTIC execution mode) keys=<k> batch=<b> win= slide= nb= lateness= par= reps=<r> (replay the stream r times, ids and timestamps advancing)
//   prints one JSON line {"mode","tuples","seconds","tuples_per_s","results","threads"}.
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <mutex>
#include <optional>
#include <string>
#include <vector>
#include <windflow.hpp>
#if defined(__CUDACC__)
#include <windflow_gpu.hpp>
#else // the CPU build (g++ -x c++): only cpu_cb, no CUDA runtime needed on the box
#define __host__
#define __device__
#endif

using namespace wf;

struct tuple64_t {
    uint64_t key, id; int64_t ivalue; double fvalue; uint64_t pad[4];
    __host__ __device__ tuple64_t(): key(0), id(0), ivalue(0), fvalue(0) { pad[0] = pad[1] = pad[2] = pad[3] = 0; }
    __host__ __device__ tuple64_t(uint64_t k, uint64_t i): key(k), id(i), ivalue(0), fvalue(0) { pad[0] = pad[1] = pad[2] = pad[3] = 0; }
};
struct result32_t {
    uint64_t key, id; int64_t isum; double fsum;
    __host__ __device__ result32_t(): key(0), id(0), isum(0), fsum(0) {}
    __host__ __device__ result32_t(uint64_t k, uint64_t i): key(k), id(i), isum(0), fsum(0) {}
};
static_assert(sizeof(tuple64_t) == 64 && sizeof(result32_t) == 32, "bench schema");

// functors of SURVEY.md 8d
struct MapF { __host__ __device__ void operator()(tuple64_t &t) { t.ivalue += 2; t.fvalue *= 1.0000001; } };
struct FiltF { __host__ __device__ bool operator()(tuple64_t &t) { return (t.ivalue & 1) == 0; } };
struct LiftF { __host__ __device__ void operator()(const tuple64_t &t, result32_t &r) { r.key = t.key; r.id = 0; r.isum = t.ivalue; r.fsum = t.fvalue; } };
struct CombF { __host__ __device__ void operator()(const result32_t &a, const result32_t &b, result32_t &o) { o.isum = a.isum + b.isum; o.fsum = a.fsum + b.fsum; } };
struct RedF { __host__ __device__ tuple64_t operator()(const tuple64_t &a, const tuple64_t &b) { tuple64_t r = a; r.ivalue = a.ivalue + b.ivalue; r.fvalue = a.fvalue + b.fvalue; return r; } };
struct KeyF { __host__ __device__ uint64_t operator()(const tuple64_t &t) { return t.key; } };
// CPU twins (the reference's CPU builders take host functors)
struct MapC { void operator()(tuple64_t &t) { t.ivalue += 2; t.fvalue *= 1.0000001; } };
struct FiltC { bool operator()(tuple64_t &t) { return (t.ivalue & 1) == 0; } };
struct LiftC { void operator()(const tuple64_t &t, result32_t &r) { r.key = t.key; r.id = 0; r.isum = t.ivalue; r.fsum = t.fvalue; } };
struct CombC { void operator()(const result32_t &a, const result32_t &b, result32_t &o) { o.isum = a.isum + b.isum; o.fsum = a.fsum + b.fsum; } };

static uint64_t splitmix64(uint64_t x) { x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; return x ^ (x >> 31); }

struct Stream { std::vector<tuple64_t> t; std::vector<uint64_t> ts, wm; };

struct SourceF {
    const Stream *s; uint64_t reps; bool set_wm;
    void operator()(Source_Shipper<tuple64_t> &sh)
    {
        const uint64_t n = s->t.size();
        const uint64_t span = n ? s->ts[n - 1] + 1 : 0;
        uint64_t max_ts = 0; // the reference refuses a watermark above the highest timestamp emitted so far (wf/source_shipper.hpp)
        for (uint64_t r = 0; r < reps; r++)
            for (uint64_t i = 0; i < n; i++) {
                if (set_wm) sh.setNextWatermark(std::min(s->wm[i] + r * span, max_ts)); // (DEFAULT mode only: the other modes have no watermarks)
                max_ts = std::max(max_ts, s->ts[i] + r * span);
                tuple64_t t = s->t[i];
                t.id += r * n;
                sh.pushWithTimestamp(std::move(t), s->ts[i] + r * span);
            }
    }
};

template <class R> struct Collected { std::mutex mu; std::vector<R> res; std::vector<uint64_t> ts; };
template <class R> struct SinkF {
    Collected<R> *c;
    void operator()(std::optional<R> &o, RuntimeContext &ctx)
    {
        if (!o) return;
        std::lock_guard<std::mutex> l(c->mu);
        c->res.push_back(*o); c->ts.push_back(ctx.getCurrentTimestamp());
    }
};

template <class R> static void dump(const std::string &path, Collected<R> &c)
{
    if (path.empty()) return;
    std::ofstream f(path, std::ios::binary);
    const uint64_t m = c.res.size();
    f.write(reinterpret_cast<const char *>(&m), 8);
    for (uint64_t i = 0; i < m; i++) { f.write(reinterpret_cast<const char *>(&c.res[i]), sizeof(R)); f.write(reinterpret_cast<const char *>(&c.ts[i]), 8); }
}

int main(int argc, char **argv)
{
    if (argc < 2) { std::fprintf(stderr, "usage: %s <gpu_cb|gpu_tb|cpu_cb|gpu_mf|gpu_red> key=value ...\n", argv[0]); return 2; }
    const std::string mode = argv[1];
    std::map<std::string, std::string> kv;
    for (int i = 2; i < argc; i++) { const char *e = std::strchr(argv[i], '='); if (e) kv[std::string(argv[i], e - argv[i])] = e + 1; }
    auto num = [&](const char *k, uint64_t d) { return kv.count(k) ? std::strtoull(kv[k].c_str(), nullptr, 10) : d; };
    const uint64_t nkeys = num("keys", 65536), batch = num("batch", 65536), win = num("win", 4096), slide = num("slide", 64), nb = num("nb", 65),
                   lateness = num("lateness", 0), par = num("par", 4), reps = num("reps", 1);
    Stream s;
    if (kv.count("in")) {
        std::ifstream f(kv["in"], std::ios::binary);
        uint64_t n = 0;
        f.read(reinterpret_cast<char *>(&n), 8);
        s.t.resize(n); s.ts.resize(n); s.wm.resize(n);
        f.read(reinterpret_cast<char *>(s.t.data()), n * 64); f.read(reinterpret_cast<char *>(s.ts.data()), n * 8); f.read(reinterpret_cast<char *>(s.wm.data()), n * 8);
        if (!f) { std::fprintf(stderr, "ref_pipeline: short read of %s\n", kv["in"].c_str()); return 2; }
    } else {
        const uint64_t n = num("gen", 1 << 20), seed = 0x5EED5EEDull;
        s.t.resize(n); s.ts.resize(n); s.wm.resize(n);
        for (uint64_t i = 0; i < n; i++) {
            tuple64_t &t = s.t[i];
            t.key = splitmix64(i) % nkeys; t.id = i;
            t.ivalue = static_cast<int64_t>(splitmix64(seed ^ i) & 0xFFFF);
            t.fvalue = static_cast<double>(splitmix64(seed ^ ~i) >> 11) * (1.0 / 9007199254740992.0);
            s.ts[i] = i; s.wm[i] = i < batch ? 0 : (i / batch) * batch - 1; // watermark of a batch = the last timestamp before it
        }
    }
    const uint64_t total = s.t.size() * reps;
    Collected<result32_t> cw; Collected<tuple64_t> ct;
    size_t threads = 0;
    const auto t0 = std::chrono::steady_clock::now();
    {
        // cpu_cb with det=1: DETERMINISTIC mode, as the reference's own count-window test (tests/win_tests/test_win_fat_cb.cpp:109): with several
        // Map/Filter replicas in front of a keyby shuffle only the ordering collectors keep a key's tuples in stream order
        const bool det = mode == "cpu_cb" && num("det", 0) != 0;
        PipeGraph graph("ref_pipeline", det ? Execution_Mode_t::DETERMINISTIC : Execution_Mode_t::DEFAULT, Time_Policy_t::EVENT_TIME);
        SourceF src{&s, reps, !det};
        if (mode == "cpu_cb") {
            Source source = Source_Builder(src).withName("source").withParallelism(1).build();
            MultiPipe &mp = graph.add_source(source);
            MapC mf; FiltC ff_;
            Map map = Map_Builder(mf).withName("map").withParallelism(par).build();
            mp.chain(map);
            Filter filter = Filter_Builder(ff_).withName("filter").withParallelism(par).build();
            mp.chain(filter);
            LiftC lf; CombC cf;
            Ffat_Windows fat = Ffat_Windows_Builder(lf, cf).withName("ffat").withParallelism(par)
                                   .withKeyBy([](const tuple64_t &t) -> uint64_t { return t.key; }).withCBWindows(win, slide).build();
            mp.add(fat);
            SinkF<result32_t> sf{&cw};
            Sink sink = Sink_Builder(sf).withName("sink").withParallelism(1).build();
            mp.chain_sink(sink);
            threads = graph.getNumThreads();
            graph.run();
        } else {
#if defined(__CUDACC__)
            Source source = Source_Builder(src).withName("source").withParallelism(1).withOutputBatchSize(batch).build();
            MultiPipe &mp = graph.add_source(source);
            if (mode == "gpu_red") {
                RedF rf;
                Reduce_GPU red = ReduceGPU_Builder(rf).withName("reduce_gpu").withParallelism(1).withKeyBy(KeyF()).build();
                mp.add(red);
                SinkF<tuple64_t> sf{&ct};
                Sink sink = Sink_Builder(sf).withName("sink").withParallelism(1).build();
                mp.chain_sink(sink);
            } else {
                MapF mf; FiltF ff_;
                Map_GPU map = MapGPU_Builder(mf).withName("map_gpu").withParallelism(1).build();
                mp.chain(map);
                Filter_GPU filter = FilterGPU_Builder(ff_).withName("filter_gpu").withParallelism(1).build();
                mp.chain(filter);
                if (mode == "gpu_mf") {
                    SinkF<tuple64_t> sf{&ct};
                    Sink sink = Sink_Builder(sf).withName("sink").withParallelism(1).build();
                    mp.chain_sink(sink);
                } else if (mode == "gpu_cb" || mode == "gpu_tb") {
                    LiftF lf; CombF cf;
                    auto b = Ffat_WindowsGPU_Builder(lf, cf).withName("ffat_gpu").withKeyBy(KeyF()).withNumWinPerBatch(nb);
                    if (mode == "gpu_cb") b.withCBWindows(win, slide);
                    else b.withTBWindows(std::chrono::microseconds(win), std::chrono::microseconds(slide)).withLateness(std::chrono::microseconds(lateness));
                    Ffat_Windows_GPU fat = b.build();
                    mp.add(fat);
                    SinkF<result32_t> sf{&cw};
                    Sink sink = Sink_Builder(sf).withName("sink").withParallelism(1).build();
                    mp.chain_sink(sink);
                } else { std::fprintf(stderr, "ref_pipeline: unknown mode %s\n", mode.c_str()); return 2; }
            }
            threads = graph.getNumThreads();
            graph.run();
#else
            std::fprintf(stderr, "ref_pipeline: this is the CPU build, mode %s needs the nvcc build\n", mode.c_str()); return 2;
#endif
        }
    }
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    const std::string out = kv.count("out") ? kv["out"] : "";
    const size_t nres = (mode == "gpu_mf" || mode == "gpu_red") ? ct.res.size() : cw.res.size();
    if (mode == "gpu_mf" || mode == "gpu_red") dump(out, ct); else dump(out, cw);
    std::printf("{\"mode\": \"%s\", \"tuples\": %llu, \"seconds\": %.6f, \"tuples_per_s\": %.1f, \"results\": %zu, \"threads\": %zu}\n", mode.c_str(),
                static_cast<unsigned long long>(total), sec, total / sec, nres, threads);
    return 0;
}
