// wf/windflow_gpu.hpp -- header-only C++17 host side of the B200-native GPU operators, keeping WindFlow's builder API
// (MapGPU_Builder / FilterGPU_Builder / ReduceGPU_Builder / Ffat_WindowsGPU_Builder, wf/builders_gpu.hpp) and the
// PipeGraph / MultiPipe wiring calls (wf/pipegraph.hpp:594-764, wf/multipipe.hpp:953-1330) for the GPU path:
//
//     PipeGraph graph("app", Execution_Mode_t::DEFAULT, Time_Policy_t::EVENT_TIME);
//     MultiPipe &mp = graph.add_source(Source_Builder(src).withOutputBatchSize(65536).build());
//     mp.chain(MapGPU_Builder(map_f).build());
//     mp.chain(FilterGPU_Builder(filter_f).build());
//     mp.add(Ffat_WindowsGPU_Builder(lift_f, comb_f).withKeyBy(key_f).withCBWindows(4096, 64).withNumWinPerBatch(65).build());
//     mp.chain_sink(Sink_Builder(sink_f).build());
//     graph.run();
//
// Every GPU replica's svc() calls libwfb200.so through the extern "C" layer of include/wfb200.h; the user's
// __host__ __device__ functors reach the kernels as a *program* registered from this translation unit
// (wfb::register_program, windflow_b200/csrc/wfb_launch.cuh). Compile the application with nvcc, link with -lwfb200.
//
// Scope of this facade: linear pipelines Source(CPU) -> GPU operators -> Sink(CPU), stateless and
// keyed-stateful Map_GPU / Filter_GPU, keyed or un-keyed Reduce_GPU, count-based and time-based Ffat_Windows_GPU, DEFAULT execution mode (the only
// mode the reference's GPU operators accept, wf/map_gpu.hpp:470-475). FastFlow is not required: stages run in the
// calling thread and hand batches over by pointer, which is what MultiPipe::chain does for chained replicas
// (wf/multipipe.hpp:538-590). Errors follow the reference convention: a red "WindFlow Error:" line and exit.
#pragma once
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <iostream>
#include <memory>
#include <optional>
#include <string>
#include <tuple>
#include <type_traits>
#include <utility>
#include <vector>
#include <cuda_runtime.h>
#include "../wfb200.h"
#include "../../windflow_b200/csrc/wfb_launch.cuh"

namespace wf {

// ---- basic types (wf/basic.hpp) ----------------------------------------------------------------------------------
enum class Execution_Mode_t { DEFAULT, DETERMINISTIC, PROBABILISTIC };
enum class Time_Policy_t { INGRESS_TIME, EVENT_TIME };
enum class Routing_Mode_t { NONE, FORWARD, KEYBY, BROADCAST, REBALANCING };
enum class Win_Type_t { CB, TB };
enum class op_type_t { SOURCE, SINK, BASIC, BASIC_GPU, WIN, WIN_PANED, WIN_MR, WIN_GPU };
struct empty_key_t {};
#define WF_RED "\033[31m"
#define WF_DEFAULT_COLOR "\033[0m"

[[noreturn]] inline void wf_fatal(const std::string &msg)
{
    std::cerr << WF_RED << "WindFlow Error: " << msg << WF_DEFAULT_COLOR << std::endl;
    std::exit(EXIT_FAILURE);
}
inline void wfbAssert(int rc, const char *file, int line) // gpuErrChk of wf/basic_gpu.hpp:99-113 for the C ABI
{
    if (rc != 0) {
        std::cerr << WF_RED << "WindFlow Error: libwfb200 => " << wfb_error_string(rc) << ", file => " << file << ", at line => " << line
                  << WF_DEFAULT_COLOR << std::endl;
        std::exit(rc > 0 ? rc : EXIT_FAILURE);
    }
}
#define wfbErrChk(ans) { ::wf::wfbAssert((ans), __FILE__, __LINE__); }
#define gpuErrChk(ans) { ::wf::wfbAssert(static_cast<int>(ans), __FILE__, __LINE__); }

// ---- functor signature sniffing (the role of wf/meta_gpu.hpp) ---------------------------------------------------------
template <class F> struct fn_sig : fn_sig<decltype(&F::operator())> {};
template <class C, class R, class... A> struct fn_sig<R (C::*)(A...) const> { using ret = R; using args = std::tuple<A...>; };
template <class C, class R, class... A> struct fn_sig<R (C::*)(A...)> { using ret = R; using args = std::tuple<A...>; };
template <class R, class... A> struct fn_sig<R (*)(A...)> { using ret = R; using args = std::tuple<A...>; };
template <class F, size_t I> using fn_arg_t = std::decay_t<std::tuple_element_t<I, typename fn_sig<F>::args>>;
template <class F> using fn_ret_t = typename fn_sig<F>::ret;

// ---- default functors of the slots an operator does not use -------------------------------------------------------------
template <class T> struct NoMap { __host__ __device__ void operator()(T &) const {} };
template <class T> struct KeepAll { __host__ __device__ bool operator()(T &) const { return true; } };
template <class T> struct NoKey { __host__ __device__ uint64_t operator()(const T &) const { return 0; } };
template <class T, class R> struct NoLift { __host__ __device__ void operator()(const T &, R &) const {} };
template <class R> struct NoComb { __host__ __device__ void operator()(const R &, const R &, R &) const {} };
template <class T> struct NoReduce { __host__ __device__ T operator()(const T &a, const T &) const { return a; } };

// The program the kernels are instantiated for: the user's functor objects travel by value in params_t.
template <class T, class R, class MapF, class FiltF, class KeyF, class LiftF, class CombF, class RedF, bool KEYED>
struct FacadeProgram {
    using tuple_t = T; using result_t = R; using key_t = uint64_t;
    struct params_t { MapF map; FiltF filt; KeyF key; LiftF lift; CombF comb; RedF red; };
    static_assert(std::is_trivially_copyable<T>::value && std::is_trivially_copyable<R>::value, "tuple_t / result_t must be trivially copyable");
    static_assert(sizeof(T) % 8 == 0 && sizeof(R) % 8 == 0, "tuple_t / result_t sizes must be multiples of 8 bytes");
    __host__ __device__ static void map(tuple_t &t, const params_t &p) { MapF f = p.map; f(t); }
    __host__ __device__ static bool filter(tuple_t &t, const params_t &p) { FiltF f = p.filt; return f(t); }
    __host__ __device__ static key_t key(const tuple_t &t, const params_t &p) { KeyF f = p.key; return static_cast<key_t>(f(t)); }
    __host__ __device__ static void lift(const tuple_t &t, result_t &r, const params_t &p) { LiftF f = p.lift; f(t, r); }
    __host__ __device__ static void comb(const result_t &a, const result_t &b, result_t &o, const params_t &p) { CombF f = p.comb; f(a, b, o); }
    __host__ __device__ static result_t make_result(key_t k, uint64_t gwid, const params_t &)
    {   // create_win_result_t_gpu, wf/basic_gpu.hpp:236-247: result_t(key, gwid) when keyed, result_t(gwid) otherwise
        if constexpr (KEYED && std::is_constructible<result_t, fn_ret_t<KeyF>, uint64_t>::value) { result_t r(static_cast<fn_ret_t<KeyF>>(k), gwid); return r; }
        else if constexpr (!KEYED && std::is_constructible<result_t, uint64_t>::value) { result_t r(gwid); return r; }
        else { (void) k; (void) gwid; return result_t(); } // programs of operators without windows never call this
    }
    __host__ __device__ static tuple_t reduce(const tuple_t &a, const tuple_t &b, const params_t &p) { RedF f = p.red; return f(a, b); }
};

// The program of a keyed-stateful Map_GPU / Filter_GPU: func(tuple_t &, state_t &) in per-key arrival order
// (API: __host__ __device__ void(tuple_t &, state_t &) / bool(tuple_t &, state_t &), wf/map_gpu.hpp:104-310, wf/filter_gpu.hpp:120-399).
template <class T, class S, class MapF2, class FiltF2, class KeyF>
struct FacadeStatefulProgram {
    using tuple_t = T; using result_t = T; using key_t = uint64_t; using state_t = S;
    struct params_t { MapF2 map; FiltF2 filt; KeyF key; };
    static_assert(std::is_trivially_copyable<T>::value && std::is_trivially_copyable<S>::value, "tuple_t / state_t must be trivially copyable");
    static_assert(sizeof(T) % 8 == 0, "tuple_t size must be a multiple of 8 bytes");
    __host__ __device__ static void map(tuple_t &, const params_t &) {}
    __host__ __device__ static bool filter(tuple_t &, const params_t &) { return true; }
    __host__ __device__ static key_t key(const tuple_t &t, const params_t &p) { KeyF f = p.key; return static_cast<key_t>(f(t)); }
    __host__ __device__ static void lift(const tuple_t &t, result_t &r, const params_t &) { r = t; }
    __host__ __device__ static void comb(const result_t &, const result_t &, result_t &, const params_t &) {}
    __host__ __device__ static result_t make_result(key_t, uint64_t, const params_t &) { return result_t(); }
    __host__ __device__ static tuple_t reduce(const tuple_t &a, const tuple_t &, const params_t &) { return a; }
    __host__ __device__ static void map_stateful(tuple_t &t, state_t &st, const params_t &p) { MapF2 f = p.map; f(t, st); }
    __host__ __device__ static bool filter_stateful(tuple_t &t, state_t &st, const params_t &p) { FiltF2 f = p.filt; return f(t, st); }
};
template <class T, class S> struct StatefulIdMap { __host__ __device__ void operator()(T &, S &) const {} };
template <class T, class S> struct StatefulKeepAll { __host__ __device__ bool operator()(T &, S &) const { return true; } };

// ---- Batch_GPU_t (wf/batch_gpu_t.hpp:50-243) as structure of arrays ---------------------------------------------------
template <class tuple_t>
struct Batch_GPU_t {
    tuple_t *tuples_gpu = nullptr;      // size * sizeof(tuple_t)
    uint64_t *ts_gpu = nullptr;         // size timestamps
    tuple_t *pinned_tuples_cpu = nullptr;
    uint64_t *pinned_ts_cpu = nullptr;
    size_t size = 0, original_size = 0;
    std::vector<uint64_t> watermarks{std::numeric_limits<uint64_t>::max()};
    bool isPunctuation = false;
    cudaStream_t cudaStream = nullptr;

    explicit Batch_GPU_t(size_t n): size(n), original_size(n)
    {
        gpuErrChk(cudaMalloc(&tuples_gpu, sizeof(tuple_t) * (n ? n : 1)));
        gpuErrChk(cudaMalloc(&ts_gpu, sizeof(uint64_t) * (n ? n : 1)));
        gpuErrChk(cudaStreamCreate(&cudaStream));
    }
    ~Batch_GPU_t()
    {
        cudaStreamSynchronize(cudaStream);
        cudaFree(tuples_gpu); cudaFree(ts_gpu);
        if (pinned_tuples_cpu) cudaFreeHost(pinned_tuples_cpu);
        if (pinned_ts_cpu) cudaFreeHost(pinned_ts_cpu);
        cudaStreamDestroy(cudaStream);
    }
    Batch_GPU_t(const Batch_GPU_t &) = delete;
    Batch_GPU_t &operator=(const Batch_GPU_t &) = delete;
    bool isPunct() const { return isPunctuation; }
    size_t getSize() const { return size; }
    uint64_t getWatermark(size_t id = 0) const { return id < watermarks.size() ? watermarks[id] : watermarks[0]; }
    void setWatermark(uint64_t wm, size_t id = 0) { if (id < watermarks.size()) watermarks[id] = wm; else watermarks[0] = wm; }
    void updateWatermark(uint64_t wm) { if (watermarks[0] > wm) watermarks[0] = wm; }
    void ensureHost()
    {
        if (!pinned_tuples_cpu) {
            gpuErrChk(cudaMallocHost(&pinned_tuples_cpu, sizeof(tuple_t) * (original_size ? original_size : 1)));
            gpuErrChk(cudaMallocHost(&pinned_ts_cpu, sizeof(uint64_t) * (original_size ? original_size : 1)));
        }
    }
    void transfer2CPU() // :154-165
    {
        ensureHost();
        gpuErrChk(cudaMemcpyAsync(pinned_tuples_cpu, tuples_gpu, sizeof(tuple_t) * size, cudaMemcpyDeviceToHost, cudaStream));
        gpuErrChk(cudaMemcpyAsync(pinned_ts_cpu, ts_gpu, sizeof(uint64_t) * size, cudaMemcpyDeviceToHost, cudaStream));
        gpuErrChk(cudaStreamSynchronize(cudaStream));
    }
    tuple_t &getTupleAtPos(size_t pos) { return pinned_tuples_cpu[pos]; }
    uint64_t getTimestampAtPos(size_t pos) { return pinned_ts_cpu[pos]; }
    void reset() { size = original_size; isPunctuation = false; watermarks.assign(1, std::numeric_limits<uint64_t>::max()); }
};

// recycling of batches (the role of wf/recycling_gpu.hpp:88-141): a free list per replica instead of an MPMC queue,
// since producer and consumer are the same thread here
template <class tuple_t>
class BatchPool {
    std::vector<Batch_GPU_t<tuple_t> *> free_;
public:
    ~BatchPool() { for (auto *b : free_) delete b; }
    Batch_GPU_t<tuple_t> *get(size_t n)
    {
        for (size_t i = 0; i < free_.size(); i++) if (free_[i]->original_size >= n) {
            auto *b = free_[i]; free_.erase(free_.begin() + i); b->reset(); b->size = n; return b;
        }
        return new Batch_GPU_t<tuple_t>(n);
    }
    void put(Batch_GPU_t<tuple_t> *b) { if (free_.size() < 8) free_.push_back(b); else delete b; }
};

// ---- stage plumbing: what FastFlow's ff_node / ff_send_out provide to chained replicas ------------------------------------
struct Stage {
    Stage *next = nullptr;
    virtual ~Stage() {}
    virtual void *svc(void *msg) = 0;     // Basic_Replica::svc, wf/basic_operator.hpp:170-195
    virtual void eosnotify() { if (next) next->eosnotify(); }
    void ff_send_out(void *msg) { if (next) next->svc(msg); }
};

class Basic_Operator {
protected:
    std::string name; size_t parallelism; Routing_Mode_t input_routing_mode; size_t outputBatchSize;
public:
    Basic_Operator(std::string n, size_t p, Routing_Mode_t r, size_t obs): name(std::move(n)), parallelism(p), input_routing_mode(r), outputBatchSize(obs) {}
    virtual ~Basic_Operator() {}
    std::string getName() const { return name; }
    size_t getParallelism() const { return parallelism; }
    Routing_Mode_t getInputRoutingMode() const { return input_routing_mode; }
    size_t getOutputBatchSize() const { return outputBatchSize; }
    virtual bool isGPUOperator() const { return true; }
    virtual std::string getType() const = 0;
    virtual std::unique_ptr<Stage> make_replica() = 0;
    void setExecutionMode(Execution_Mode_t m) { if (m != Execution_Mode_t::DEFAULT) wf_fatal(getType() + " can only be used in DEFAULT mode"); }
};

// ---- Source / Sink (CPU side) -----------------------------------------------------------------------------------------
template <class tuple_t> class Source_Shipper;

template <class tuple_t>
class SourceStage: public Stage { // Source_Replica + Forward_Emitter_GPU<..., false, true> (wf/forward_emitter_gpu.hpp:254-305)
    friend class Source_Shipper<tuple_t>;
    std::function<void(Source_Shipper<tuple_t> &)> func;
    size_t batch_size;
    BatchPool<tuple_t> pool;
    Batch_GPU_t<tuple_t> *cur = nullptr;
    size_t fill = 0;
    uint64_t next_wm = 0;
public:
    SourceStage(std::function<void(Source_Shipper<tuple_t> &)> f, size_t bs): func(std::move(f)), batch_size(bs) {}
    void *svc(void *) override { return nullptr; }
    void push(const tuple_t &t, uint64_t ts)
    {
        if (!cur) { cur = pool.get(batch_size); cur->ensureHost(); fill = 0; }
        cur->pinned_tuples_cpu[fill] = t; cur->pinned_ts_cpu[fill] = ts; cur->updateWatermark(next_wm);
        if (++fill == batch_size) flush();
    }
    void flush()
    {
        if (!cur || fill == 0) return;
        cur->size = fill;
        gpuErrChk(cudaMemcpyAsync(cur->tuples_gpu, cur->pinned_tuples_cpu, sizeof(tuple_t) * fill, cudaMemcpyHostToDevice, cur->cudaStream));
        gpuErrChk(cudaMemcpyAsync(cur->ts_gpu, cur->pinned_ts_cpu, sizeof(uint64_t) * fill, cudaMemcpyHostToDevice, cur->cudaStream));
        Batch_GPU_t<tuple_t> *b = cur; cur = nullptr; fill = 0;
        this->ff_send_out(b); // ownership moves downstream; the last GPU stage / the sink returns it through recycle()
    }
    void recycle(Batch_GPU_t<tuple_t> *b) { pool.put(b); }
    void run();
};

template <class tuple_t>
class Source_Shipper { // wf/source_shipper.hpp:289-322
    SourceStage<tuple_t> *st;
public:
    explicit Source_Shipper(SourceStage<tuple_t> *s): st(s) {}
    void push(const tuple_t &t) { st->push(t, std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now().time_since_epoch()).count()); }
    void pushWithTimestamp(const tuple_t &t, uint64_t ts) { st->push(t, ts); }
    void setNextWatermark(uint64_t wm) { st->next_wm = wm; }
};
template <class tuple_t> void SourceStage<tuple_t>::run() { Source_Shipper<tuple_t> sh(this); func(sh); flush(); this->eosnotify(); }

template <class source_func_t>
class Source: public Basic_Operator {
public:
    source_func_t func;
    static constexpr op_type_t op_type = op_type_t::SOURCE;
    Source(source_func_t f, std::string n, size_t p, size_t obs): Basic_Operator(std::move(n), p, Routing_Mode_t::NONE, obs), func(f) {}
    bool isGPUOperator() const override { return false; }
    std::string getType() const override { return "Source"; }
    std::unique_ptr<Stage> make_replica() override { return nullptr; }
};
template <class F> struct shipper_tuple;
template <class T> struct shipper_tuple<Source_Shipper<T>> { using type = T; };

template <class source_func_t>
class Source_Builder {
    source_func_t func; std::string name = "source"; size_t parallelism = 1, obs = 0;
public:
    explicit Source_Builder(source_func_t f): func(f) {}
    auto &withName(std::string n) { name = std::move(n); return *this; }
    auto &withParallelism(size_t p) { parallelism = p; return *this; }
    auto &withOutputBatchSize(size_t n) { obs = n; return *this; }
    auto build() { return Source<source_func_t>(func, name, parallelism, obs); }
};

template <class sink_func_t>
class Sink: public Basic_Operator {
public:
    sink_func_t func;
    static constexpr op_type_t op_type = op_type_t::SINK;
    Sink(sink_func_t f, std::string n, size_t p): Basic_Operator(std::move(n), p, Routing_Mode_t::FORWARD, 0), func(f) {}
    bool isGPUOperator() const override { return false; }
    std::string getType() const override { return "Sink"; }
    std::unique_ptr<Stage> make_replica() override { return nullptr; }
};
template <class sink_func_t>
class Sink_Builder {
    sink_func_t func; std::string name = "sink"; size_t parallelism = 1;
public:
    explicit Sink_Builder(sink_func_t f): func(f) {}
    auto &withName(std::string n) { name = std::move(n); return *this; }
    auto &withParallelism(size_t p) { parallelism = p; return *this; }
    auto build() { return Sink<sink_func_t>(func, name, parallelism); }
};

template <class tuple_t, class sink_func_t>
class SinkStage: public Stage { // Forward_Emitter_GPU<..., true, false> (transfer2CPU) + Sink_Replica::svc (wf/sink.hpp:102-111)
    sink_func_t func;
    std::function<void(void *)> recycle;
public:
    SinkStage(sink_func_t f, std::function<void(void *)> r): func(f), recycle(std::move(r)) {}
    void *svc(void *msg) override
    {
        auto *b = reinterpret_cast<Batch_GPU_t<tuple_t> *>(msg);
        if (!b->isPunct() && b->size) {
            b->transfer2CPU();
            for (size_t i = 0; i < b->size; i++) { std::optional<tuple_t> o(b->getTupleAtPos(i)); func(o); }
        }
        recycle(b);
        return nullptr;
    }
    void eosnotify() override { std::optional<tuple_t> o; func(o); } // the reference's end-of-stream call with an empty optional
};

// ---- GPU operators ------------------------------------------------------------------------------------------------------
// Map_GPU, stateless (wf/map_gpu.hpp:313-420). svc: wfb_map in place on the batch's own stream.
template <class map_func_gpu_t>
class Map_GPU: public Basic_Operator {
public:
    using tuple_t = fn_arg_t<map_func_gpu_t, 0>;
    using result_t = tuple_t;
    using prog_t = FacadeProgram<tuple_t, tuple_t, map_func_gpu_t, KeepAll<tuple_t>, NoKey<tuple_t>, NoLift<tuple_t, tuple_t>, NoComb<tuple_t>, NoReduce<tuple_t>, false>;
    static constexpr op_type_t op_type = op_type_t::BASIC_GPU;
    map_func_gpu_t func;
    Map_GPU(map_func_gpu_t f, size_t p, std::string n, Routing_Mode_t r): Basic_Operator(std::move(n), p, r, 1), func(f) {}
    std::string getType() const override { return "Map_GPU"; }
    struct Replica: Stage {
        wfb_engine_t *eng = nullptr; typename prog_t::params_t prm;
        explicit Replica(map_func_gpu_t f): prm{f, {}, {}, {}, {}, {}} { wfbErrChk(wfb_engine_create(&eng, wfb::register_program<prog_t>())); }
        ~Replica() override { wfb_engine_destroy(eng); }
        void *svc(void *msg) override
        {
            auto *in = reinterpret_cast<Batch_GPU_t<tuple_t> *>(msg);
            if (!in->isPunct()) wfbErrChk(wfb_map(eng, reinterpret_cast<const wfb_functors_t *>(&prm), in->tuples_gpu, static_cast<uint32_t>(in->size), in->cudaStream));
            this->ff_send_out(in); // doEmit_inplace: the same batch moves on
            return nullptr;
        }
    };
    std::unique_ptr<Stage> make_replica() override { return std::make_unique<Replica>(func); }
};

// Filter_GPU, stateless (wf/filter_gpu.hpp:401-600). svc: wfb_map_filter into a spare batch, then the spare moves on.
template <class filter_func_gpu_t>
class Filter_GPU: public Basic_Operator {
public:
    using tuple_t = fn_arg_t<filter_func_gpu_t, 0>;
    using result_t = tuple_t;
    using prog_t = FacadeProgram<tuple_t, tuple_t, NoMap<tuple_t>, filter_func_gpu_t, NoKey<tuple_t>, NoLift<tuple_t, tuple_t>, NoComb<tuple_t>, NoReduce<tuple_t>, false>;
    static constexpr op_type_t op_type = op_type_t::BASIC_GPU;
    filter_func_gpu_t func;
    std::function<void(void *)> recycle_in; // set by MultiPipe: returns consumed input batches to the source's pool
    Filter_GPU(filter_func_gpu_t f, size_t p, std::string n, Routing_Mode_t r): Basic_Operator(std::move(n), p, r, 1), func(f) {}
    std::string getType() const override { return "Filter_GPU"; }
    struct Replica: Stage {
        wfb_engine_t *eng = nullptr; typename prog_t::params_t prm; uint32_t *n_out_dev = nullptr; uint32_t *n_out_h = nullptr;
        BatchPool<tuple_t> pool; std::function<void(void *)> recycle_in;
        Replica(filter_func_gpu_t f, std::function<void(void *)> r): prm{{}, f, {}, {}, {}, {}}, recycle_in(std::move(r))
        {
            wfbErrChk(wfb_engine_create(&eng, wfb::register_program<prog_t>()));
            gpuErrChk(cudaMalloc(&n_out_dev, sizeof(uint32_t))); gpuErrChk(cudaMallocHost(&n_out_h, sizeof(uint32_t)));
        }
        ~Replica() override { wfb_engine_destroy(eng); cudaFree(n_out_dev); cudaFreeHost(n_out_h); }
        void *svc(void *msg) override
        {
            auto *in = reinterpret_cast<Batch_GPU_t<tuple_t> *>(msg);
            if (in->isPunct()) { this->ff_send_out(in); return nullptr; }
            Batch_GPU_t<tuple_t> *out = pool.get(in->original_size);
            out->watermarks = in->watermarks;
            wfbErrChk(wfb_map_filter(eng, reinterpret_cast<const wfb_functors_t *>(&prm), in->tuples_gpu, in->ts_gpu, static_cast<uint32_t>(in->size),
                                     out->tuples_gpu, out->ts_gpu, n_out_dev, in->cudaStream));
            gpuErrChk(cudaMemcpyAsync(n_out_h, n_out_dev, sizeof(uint32_t), cudaMemcpyDeviceToHost, in->cudaStream));
            gpuErrChk(cudaStreamSynchronize(in->cudaStream)); // the reference syncs here too (wf/filter_gpu.hpp:571)
            out->size = *n_out_h;
            if (recycle_in) recycle_in(in); else delete in;
            if (out->size == 0) { pool.put(out); return nullptr; } // empty batch dropped (:572-581)
            this->ff_send_out(out);
            return nullptr;
        }
        void take_back(void *b) { pool.put(reinterpret_cast<Batch_GPU_t<tuple_t> *>(b)); }
    };
    std::unique_ptr<Stage> make_replica() override { return std::make_unique<Replica>(func, recycle_in); }
};

// Map_GPU, keyed-stateful (wf/map_gpu.hpp:104-310): func(tuple, state_of_key) in per-key arrival order, in place. The key -> state
// table is one wfb_kstate_t per operator, shared by its replicas (the reference's TBB map + spinlock, :551-559).
template <class map_func_gpu_t, class keyextr_func_gpu_t>
class Map_GPU_KB: public Basic_Operator {
public:
    using tuple_t = fn_arg_t<map_func_gpu_t, 0>;
    using state_t = fn_arg_t<map_func_gpu_t, 1>;
    using result_t = tuple_t;
    using prog_t = FacadeStatefulProgram<tuple_t, state_t, map_func_gpu_t, StatefulKeepAll<tuple_t, state_t>, keyextr_func_gpu_t>;
    static constexpr op_type_t op_type = op_type_t::BASIC_GPU;
    map_func_gpu_t func; keyextr_func_gpu_t key_extr; uint32_t max_keys;
    std::shared_ptr<wfb_kstate_t> kstate;
    Map_GPU_KB(map_func_gpu_t f, keyextr_func_gpu_t k, size_t p, std::string n, uint32_t mk): Basic_Operator(std::move(n), p, Routing_Mode_t::KEYBY, 1), func(f), key_extr(k), max_keys(mk) {}
    std::string getType() const override { return "Map_GPU"; }
    keyextr_func_gpu_t getKeyExtractor() const { return key_extr; }
    struct Replica: Stage {
        std::shared_ptr<wfb_kstate_t> ks; typename prog_t::params_t prm;
        Replica(std::shared_ptr<wfb_kstate_t> h, map_func_gpu_t f, keyextr_func_gpu_t k): ks(std::move(h)), prm{f, {}, k} {}
        void *svc(void *msg) override
        {
            auto *in = reinterpret_cast<Batch_GPU_t<tuple_t> *>(msg);
            if (!in->isPunct() && in->size) {
                wfb_batch_t b{}; b.tuples = in->tuples_gpu; b.ts = in->ts_gpu; b.n = static_cast<uint32_t>(in->size);
                wfbErrChk(wfb_map_stateful(ks.get(), reinterpret_cast<const wfb_functors_t *>(&prm), &b, 1, in->cudaStream));
            }
            this->ff_send_out(in);
            return nullptr;
        }
    };
    std::unique_ptr<Stage> make_replica() override
    {
        if (!kstate) { wfb_kstate_t *h = nullptr; wfbErrChk(wfb_kstate_create(&h, wfb::register_program<prog_t>(), max_keys, 0)); kstate.reset(h, [](wfb_kstate_t *p) { wfb_kstate_destroy(p); }); }
        return std::make_unique<Replica>(kstate, func, key_extr);
    }
};

// Filter_GPU, keyed-stateful (wf/filter_gpu.hpp:120-399): predicate(tuple, state_of_key); survivors compacted (stable).
template <class filter_func_gpu_t, class keyextr_func_gpu_t>
class Filter_GPU_KB: public Basic_Operator {
public:
    using tuple_t = fn_arg_t<filter_func_gpu_t, 0>;
    using state_t = fn_arg_t<filter_func_gpu_t, 1>;
    using result_t = tuple_t;
    using prog_t = FacadeStatefulProgram<tuple_t, state_t, StatefulIdMap<tuple_t, state_t>, filter_func_gpu_t, keyextr_func_gpu_t>;
    static constexpr op_type_t op_type = op_type_t::BASIC_GPU;
    filter_func_gpu_t func; keyextr_func_gpu_t key_extr; uint32_t max_keys;
    std::shared_ptr<wfb_kstate_t> kstate;
    std::function<void(void *)> recycle_in;
    Filter_GPU_KB(filter_func_gpu_t f, keyextr_func_gpu_t k, size_t p, std::string n, uint32_t mk): Basic_Operator(std::move(n), p, Routing_Mode_t::KEYBY, 1), func(f), key_extr(k), max_keys(mk) {}
    std::string getType() const override { return "Filter_GPU"; }
    keyextr_func_gpu_t getKeyExtractor() const { return key_extr; }
    struct Replica: Stage {
        std::shared_ptr<wfb_kstate_t> ks; typename prog_t::params_t prm; uint32_t *n_out_dev = nullptr; uint32_t *n_out_h = nullptr;
        BatchPool<tuple_t> pool; std::function<void(void *)> recycle_in;
        Replica(std::shared_ptr<wfb_kstate_t> h, filter_func_gpu_t f, keyextr_func_gpu_t k, std::function<void(void *)> r): ks(std::move(h)), prm{{}, f, k}, recycle_in(std::move(r))
        {
            gpuErrChk(cudaMalloc(&n_out_dev, sizeof(uint32_t))); gpuErrChk(cudaMallocHost(&n_out_h, sizeof(uint32_t)));
        }
        ~Replica() override { cudaFree(n_out_dev); cudaFreeHost(n_out_h); }
        void *svc(void *msg) override
        {
            auto *in = reinterpret_cast<Batch_GPU_t<tuple_t> *>(msg);
            if (in->isPunct()) { this->ff_send_out(in); return nullptr; }
            Batch_GPU_t<tuple_t> *out = pool.get(in->original_size);
            out->watermarks = in->watermarks;
            wfb_batch_t bi{}, bo{};
            bi.tuples = in->tuples_gpu; bi.ts = in->ts_gpu; bi.n = static_cast<uint32_t>(in->size);
            bo.tuples = out->tuples_gpu; bo.ts = out->ts_gpu; bo.n = bi.n;
            wfbErrChk(wfb_filter_stateful(ks.get(), reinterpret_cast<const wfb_functors_t *>(&prm), &bi, &bo, 1, n_out_dev, in->cudaStream));
            gpuErrChk(cudaMemcpyAsync(n_out_h, n_out_dev, sizeof(uint32_t), cudaMemcpyDeviceToHost, in->cudaStream));
            gpuErrChk(cudaStreamSynchronize(in->cudaStream));
            out->size = *n_out_h;
            if (recycle_in) recycle_in(in); else delete in;
            if (out->size == 0) { pool.put(out); return nullptr; }
            this->ff_send_out(out);
            return nullptr;
        }
        void take_back(void *b) { pool.put(reinterpret_cast<Batch_GPU_t<tuple_t> *>(b)); }
    };
    std::unique_ptr<Stage> make_replica() override
    {
        if (!kstate) { wfb_kstate_t *h = nullptr; wfbErrChk(wfb_kstate_create(&h, wfb::register_program<prog_t>(), max_keys, 0)); kstate.reset(h, [](wfb_kstate_t *p) { wfb_kstate_destroy(p); }); }
        return std::make_unique<Replica>(kstate, func, key_extr, recycle_in);
    }
};

// Reduce_GPU (wf/reduce_gpu.hpp:109-289): per batch, one item per distinct key (ascending) or one item for the batch.
template <class reduce_func_gpu_t, class keyextr_func_gpu_t>
class Reduce_GPU: public Basic_Operator {
public:
    using tuple_t = fn_arg_t<reduce_func_gpu_t, 0>;
    using result_t = tuple_t;
    static constexpr bool isKeyed = !std::is_same<keyextr_func_gpu_t, NoKey<tuple_t>>::value;
    using prog_t = FacadeProgram<tuple_t, tuple_t, NoMap<tuple_t>, KeepAll<tuple_t>, keyextr_func_gpu_t, NoLift<tuple_t, tuple_t>, NoComb<tuple_t>, reduce_func_gpu_t, false>;
    static constexpr op_type_t op_type = op_type_t::BASIC_GPU;
    reduce_func_gpu_t func; keyextr_func_gpu_t key_extr;
    std::function<void(void *)> recycle_in;
    Reduce_GPU(reduce_func_gpu_t f, keyextr_func_gpu_t k, size_t p, std::string n, Routing_Mode_t r): Basic_Operator(std::move(n), p, r, 1), func(f), key_extr(k) {}
    std::string getType() const override { return "Reduce_GPU"; }
    struct Replica: Stage {
        wfb_engine_t *eng = nullptr; typename prog_t::params_t prm; uint32_t *n_out_dev = nullptr; uint32_t *n_out_h = nullptr;
        BatchPool<tuple_t> pool; std::function<void(void *)> recycle_in;
        Replica(reduce_func_gpu_t f, keyextr_func_gpu_t k, std::function<void(void *)> r): prm{{}, {}, k, {}, {}, f}, recycle_in(std::move(r))
        {
            wfbErrChk(wfb_engine_create(&eng, wfb::register_program<prog_t>()));
            wfbErrChk(wfb_engine_set_params(eng, &prm, sizeof(prm)));
            gpuErrChk(cudaMalloc(&n_out_dev, sizeof(uint32_t))); gpuErrChk(cudaMallocHost(&n_out_h, sizeof(uint32_t)));
        }
        ~Replica() override { wfb_engine_destroy(eng); cudaFree(n_out_dev); cudaFreeHost(n_out_h); }
        void *svc(void *msg) override
        {
            auto *in = reinterpret_cast<Batch_GPU_t<tuple_t> *>(msg);
            if (in->isPunct()) { this->ff_send_out(in); return nullptr; }
            Batch_GPU_t<tuple_t> *out = pool.get(in->original_size);
            out->watermarks = in->watermarks;
            if constexpr (isKeyed) {
                wfbErrChk(wfb_reduce_by_key(eng, in->tuples_gpu, in->ts_gpu, static_cast<uint32_t>(in->size), out->tuples_gpu, out->ts_gpu, n_out_dev, in->cudaStream));
                gpuErrChk(cudaMemcpyAsync(n_out_h, n_out_dev, sizeof(uint32_t), cudaMemcpyDeviceToHost, in->cudaStream));
                gpuErrChk(cudaStreamSynchronize(in->cudaStream));
                out->size = *n_out_h;
            } else {
                wfbErrChk(wfb_reduce_all(eng, in->tuples_gpu, in->ts_gpu, static_cast<uint32_t>(in->size), out->tuples_gpu, out->ts_gpu, in->cudaStream));
                gpuErrChk(cudaStreamSynchronize(in->cudaStream));
                out->size = 1;
            }
            if (recycle_in) recycle_in(in); else delete in;
            this->ff_send_out(out);
            return nullptr;
        }
    };
    std::unique_ptr<Stage> make_replica() override { return std::make_unique<Replica>(func, key_extr, recycle_in); }
};

// Ffat_Windows_GPU, count-based (wf/ffat_windows_gpu.hpp:59-274, wf/ffat_replica_gpu.hpp:707-867)
template <class lift_func_gpu_t, class comb_func_gpu_t, class keyextr_func_gpu_t>
class Ffat_Windows_GPU: public Basic_Operator {
public:
    using tuple_t = fn_arg_t<lift_func_gpu_t, 0>;
    using result_t = fn_arg_t<lift_func_gpu_t, 1>;
    static constexpr bool isKeyed = !std::is_same<keyextr_func_gpu_t, NoKey<tuple_t>>::value;
    using prog_t = FacadeProgram<tuple_t, result_t, NoMap<tuple_t>, KeepAll<tuple_t>, keyextr_func_gpu_t, lift_func_gpu_t, comb_func_gpu_t, NoReduce<tuple_t>, isKeyed>;
    static constexpr op_type_t op_type = op_type_t::WIN_GPU;
    lift_func_gpu_t lift; comb_func_gpu_t comb; keyextr_func_gpu_t key_extr;
    uint64_t win_len, slide_len, lateness; Win_Type_t winType; size_t numWinPerBatch; uint32_t max_keys;
    std::function<void(void *)> recycle_in;
    Ffat_Windows_GPU(lift_func_gpu_t l, comb_func_gpu_t c, keyextr_func_gpu_t k, std::string n, uint64_t w, uint64_t s, uint64_t late,
                     Win_Type_t wt, size_t nwb, uint32_t mk):
        Basic_Operator(std::move(n), 1 /* forced to 1, wf/ffat_windows_gpu.hpp:197 */, Routing_Mode_t::FORWARD, nwb), lift(l), comb(c), key_extr(k),
        win_len(w), slide_len(s), lateness(late), winType(wt), numWinPerBatch(nwb), max_keys(mk)
    {
        if (win_len == 0 || slide_len == 0) wf_fatal("Ffat_Windows_GPU used with window length or slide equal to zero");
        if (numWinPerBatch == 0) wf_fatal("Ffat_Windows_GPU used with zero windows per batch");
    }
    std::string getType() const override { return "Ffat_Windows_GPU"; }
    struct Replica: Stage {
        wfb_ffat_t *ffat = nullptr; typename prog_t::params_t prm; uint32_t *n_out_dev = nullptr; uint32_t *n_out_h = nullptr;
        BatchPool<result_t> pool; std::function<void(void *)> recycle_in; uint64_t slide, nb; uint32_t max_keys; bool tb; uint64_t last_wm = 0;
        Replica(const Ffat_Windows_GPU &op): prm{{}, {}, op.key_extr, op.lift, op.comb, {}}, recycle_in(op.recycle_in), slide(op.slide_len),
                                             nb(op.numWinPerBatch), max_keys(op.max_keys), tb(op.winType == Win_Type_t::TB)
        {
            wfbErrChk(wfb_ffat_create(&ffat, wfb::register_program<prog_t>(), op.win_len, op.slide_len, static_cast<uint32_t>(op.numWinPerBatch),
                                      op.max_keys, tb ? 1 : 0, op.lateness, 0));
            wfbErrChk(wfb_ffat_set_params(ffat, &prm, sizeof(prm)));
            gpuErrChk(cudaMalloc(&n_out_dev, sizeof(uint32_t))); gpuErrChk(cudaMallocHost(&n_out_h, sizeof(uint32_t)));
        }
        ~Replica() override { wfb_ffat_destroy(ffat); cudaFree(n_out_dev); cudaFreeHost(n_out_h); }
        void *svc(void *msg) override
        {
            auto *in = reinterpret_cast<Batch_GPU_t<tuple_t> *>(msg);
            if (in->isPunct()) { if (recycle_in) recycle_in(in); else delete in; return nullptr; }
            // every group that can fire on this batch: per key, count-based one per slide*nb items; time-based one per slide*nb
            // time units the watermark advanced (+1: the first group, B panes, may complete together with the next one)
            const uint64_t wm = in->getWatermark();
            const size_t per_key = tb ? static_cast<size_t>((wm > last_wm ? wm - last_wm : 0) / (slide * nb) + 2) : 1;
            const size_t cap = tb ? static_cast<size_t>(max_keys) * per_key * nb : (in->size / (slide * nb) + max_keys + 1) * nb;
            if (tb) last_wm = wm;
            Batch_GPU_t<result_t> *out = pool.get(cap);
            wfb_batch_t b{in->tuples_gpu, in->ts_gpu, wm, static_cast<uint32_t>(in->size), 0};
            if (tb) wfbErrChk(wfb_ffat_process_tb(ffat, nullptr, &b, 1, out->tuples_gpu, out->ts_gpu, static_cast<uint32_t>(cap), n_out_dev, in->cudaStream))
            else wfbErrChk(wfb_ffat_process_cb(ffat, nullptr, &b, 1, out->tuples_gpu, out->ts_gpu, static_cast<uint32_t>(cap), n_out_dev, in->cudaStream));
            gpuErrChk(cudaMemcpyAsync(n_out_h, n_out_dev, sizeof(uint32_t), cudaMemcpyDeviceToHost, in->cudaStream));
            gpuErrChk(cudaStreamSynchronize(in->cudaStream));
            out->size = *n_out_h; out->setWatermark(in->getWatermark());
            if (recycle_in) recycle_in(in); else delete in;
            if (out->size == 0) { pool.put(out); return nullptr; }
            // results were produced on the input batch's stream and are complete (synchronised above)
            this->ff_send_out(out);
            return nullptr;
        }
    };
    std::unique_ptr<Stage> make_replica() override { return std::make_unique<Replica>(*this); }
};

// ---- builders (wf/builders_gpu.hpp) ----------------------------------------------------------------------------------------
template <class map_func_gpu_t, class keyextr_func_gpu_t>
class MapGPU_KB_Builder { // MapGPU_Builder(func).withKeyBy(key_extr): the keyed-stateful operator
    map_func_gpu_t func; keyextr_func_gpu_t key; std::string name; size_t parallelism; uint32_t max_keys = 1u << 16;
public:
    MapGPU_KB_Builder(map_func_gpu_t f, keyextr_func_gpu_t k, std::string n, size_t p): func(f), key(k), name(std::move(n)), parallelism(p) {}
    auto &withName(std::string n) { name = std::move(n); return *this; }
    auto &withParallelism(size_t p) { parallelism = p; return *this; }
    auto &withMaxKeys(uint32_t mk) { max_keys = mk; return *this; } // capacity of the device key -> state table (not in the reference: its map grows on the host)
    auto build() { return Map_GPU_KB<map_func_gpu_t, keyextr_func_gpu_t>(func, key, parallelism, name, max_keys); }
};

template <class map_func_gpu_t>
class MapGPU_Builder {
    map_func_gpu_t func; std::string name = "map_gpu"; size_t parallelism = 1; Routing_Mode_t mode = Routing_Mode_t::FORWARD;
    static constexpr size_t arity = std::tuple_size<typename fn_sig<map_func_gpu_t>::args>::value;
public:
    explicit MapGPU_Builder(map_func_gpu_t f): func(f)
    {
        static_assert(arity == 1 || arity == 2,
                      "WindFlow Compilation Error - MapGPU_Builder: __host__ __device__ void(tuple_t &) or void(tuple_t &, state_t &)");
    }
    auto &withName(std::string n) { name = std::move(n); return *this; }
    auto &withParallelism(size_t p) { parallelism = p; return *this; }
    auto &withRebalancing() { mode = Routing_Mode_t::REBALANCING; return *this; }
    template <class keyextr_t> auto withKeyBy(keyextr_t k)
    {
        static_assert(arity == 2, "WindFlow Compilation Error - MapGPU_Builder: withKeyBy() needs the stateful signature void(tuple_t &, state_t &)");
        return MapGPU_KB_Builder<map_func_gpu_t, keyextr_t>(func, k, name, parallelism);
    }
    auto build()
    {
        static_assert(arity == 1, "WindFlow Compilation Error - MapGPU_Builder: a stateful functor needs withKeyBy()");
        return Map_GPU<map_func_gpu_t>(func, parallelism, name, mode);
    }
};

template <class filter_func_gpu_t, class keyextr_func_gpu_t>
class FilterGPU_KB_Builder {
    filter_func_gpu_t func; keyextr_func_gpu_t key; std::string name; size_t parallelism; uint32_t max_keys = 1u << 16;
public:
    FilterGPU_KB_Builder(filter_func_gpu_t f, keyextr_func_gpu_t k, std::string n, size_t p): func(f), key(k), name(std::move(n)), parallelism(p) {}
    auto &withName(std::string n) { name = std::move(n); return *this; }
    auto &withParallelism(size_t p) { parallelism = p; return *this; }
    auto &withMaxKeys(uint32_t mk) { max_keys = mk; return *this; }
    auto build() { return Filter_GPU_KB<filter_func_gpu_t, keyextr_func_gpu_t>(func, key, parallelism, name, max_keys); }
};

template <class filter_func_gpu_t>
class FilterGPU_Builder {
    filter_func_gpu_t func; std::string name = "filter_gpu"; size_t parallelism = 1; Routing_Mode_t mode = Routing_Mode_t::FORWARD;
    static constexpr size_t arity = std::tuple_size<typename fn_sig<filter_func_gpu_t>::args>::value;
public:
    explicit FilterGPU_Builder(filter_func_gpu_t f): func(f)
    {
        static_assert(arity == 1 || arity == 2,
                      "WindFlow Compilation Error - FilterGPU_Builder: __host__ __device__ bool(tuple_t &) or bool(tuple_t &, state_t &)");
    }
    auto &withName(std::string n) { name = std::move(n); return *this; }
    auto &withParallelism(size_t p) { parallelism = p; return *this; }
    auto &withRebalancing() { mode = Routing_Mode_t::REBALANCING; return *this; }
    template <class keyextr_t> auto withKeyBy(keyextr_t k)
    {
        static_assert(arity == 2, "WindFlow Compilation Error - FilterGPU_Builder: withKeyBy() needs the stateful signature bool(tuple_t &, state_t &)");
        return FilterGPU_KB_Builder<filter_func_gpu_t, keyextr_t>(func, k, name, parallelism);
    }
    auto build()
    {
        static_assert(arity == 1, "WindFlow Compilation Error - FilterGPU_Builder: a stateful functor needs withKeyBy()");
        return Filter_GPU<filter_func_gpu_t>(func, parallelism, name, mode);
    }
};

template <class reduce_func_gpu_t, class keyextr_func_gpu_t = NoKey<fn_arg_t<reduce_func_gpu_t, 0>>>
class ReduceGPU_Builder {
    template <class A, class B> friend class ReduceGPU_Builder;
    reduce_func_gpu_t func; keyextr_func_gpu_t key_extr; std::string name = "reduce_gpu"; size_t parallelism = 1;
    Routing_Mode_t mode = Routing_Mode_t::FORWARD;
    ReduceGPU_Builder(reduce_func_gpu_t f, keyextr_func_gpu_t k): func(f), key_extr(k) {}
public:
    explicit ReduceGPU_Builder(reduce_func_gpu_t f): func(f), key_extr() {}
    auto &withName(std::string n) { name = std::move(n); return *this; }
    auto &withParallelism(size_t p) { parallelism = p; return *this; }
    template <class new_keyextr_t> auto withKeyBy(new_keyextr_t k)
    {
        ReduceGPU_Builder<reduce_func_gpu_t, new_keyextr_t> nb(func, k);
        nb.name = name; nb.parallelism = parallelism; nb.mode = Routing_Mode_t::KEYBY;
        return nb;
    }
    auto build() { return Reduce_GPU<reduce_func_gpu_t, keyextr_func_gpu_t>(func, key_extr, parallelism, name, mode); }
};

template <class lift_func_gpu_t, class comb_func_gpu_t, class keyextr_func_gpu_t = NoKey<fn_arg_t<lift_func_gpu_t, 0>>>
class Ffat_WindowsGPU_Builder {
    template <class A, class B, class C> friend class Ffat_WindowsGPU_Builder;
    lift_func_gpu_t lift; comb_func_gpu_t comb; keyextr_func_gpu_t key_extr; std::string name = "ffat_windows_gpu";
    size_t numWinPerBatch = 0; uint64_t win_len = 0, slide_len = 0, lateness = 0; Win_Type_t winType = Win_Type_t::CB; uint32_t max_keys = 65536;
    Ffat_WindowsGPU_Builder(lift_func_gpu_t l, comb_func_gpu_t c, keyextr_func_gpu_t k): lift(l), comb(c), key_extr(k) {}
public:
    Ffat_WindowsGPU_Builder(lift_func_gpu_t l, comb_func_gpu_t c): lift(l), comb(c), key_extr() {}
    auto &withName(std::string n) { name = std::move(n); return *this; }
    template <class new_keyextr_t> auto withKeyBy(new_keyextr_t k)
    {
        Ffat_WindowsGPU_Builder<lift_func_gpu_t, comb_func_gpu_t, new_keyextr_t> nb(lift, comb, k);
        nb.name = name; nb.numWinPerBatch = numWinPerBatch; nb.win_len = win_len; nb.slide_len = slide_len; nb.lateness = lateness;
        nb.winType = winType; nb.max_keys = max_keys;
        return nb;
    }
    auto &withCBWindows(uint64_t w, uint64_t s) { win_len = w; slide_len = s; winType = Win_Type_t::CB; lateness = 0; return *this; }
    auto &withTBWindows(std::chrono::microseconds w, std::chrono::microseconds s) { win_len = w.count(); slide_len = s.count(); winType = Win_Type_t::TB; return *this; }
    auto &withLateness(std::chrono::microseconds l) { lateness = l.count(); return *this; }
    auto &withNumWinPerBatch(size_t n) { numWinPerBatch = n; return *this; }
    auto &withMaxKeys(uint32_t n) { max_keys = n; return *this; } // extension: capacity of the device-resident key table
    auto build() { return Ffat_Windows_GPU<lift_func_gpu_t, comb_func_gpu_t, keyextr_func_gpu_t>(lift, comb, key_extr, name, win_len, slide_len, lateness, winType, numWinPerBatch, max_keys); }
};

// ---- MultiPipe / PipeGraph for linear GPU pipelines (wf/multipipe.hpp, wf/pipegraph.hpp) -----------------------------------
class PipeGraph;
class MultiPipe {
    friend class PipeGraph;
    std::vector<std::unique_ptr<Stage>> stages;
    std::vector<std::string> op_names;
    std::function<void()> run_source;
    std::function<void(void *)> recycle_prev; // returns a batch of the CURRENT tail's output type to its producer's pool
    bool has_sink = false; size_t prevOutputBatchSize = 0; bool tail_is_gpu = false;
    void link(std::unique_ptr<Stage> st) { if (!stages.empty()) stages.back()->next = st.get(); stages.push_back(std::move(st)); }
    template <class op_t> void attach(op_t &op)
    {
        if (has_sink) wf_fatal("MultiPipe is already terminated by a Sink");
        if (!tail_is_gpu && prevOutputBatchSize == 0)  // wf/multipipe.hpp:486-488
            wf_fatal(op.getType() + " cannot be added after a CPU operator without withOutputBatchSize()");
        op.setExecutionMode(Execution_Mode_t::DEFAULT);
        op_names.push_back(op.getName());
    }
public:
    template <class op_t> MultiPipe &add(op_t op) { return chain(op); } // single thread: add == chain
    template <class map_f> MultiPipe &chain(Map_GPU<map_f> op)
    {
        attach(op); link(op.make_replica()); tail_is_gpu = true; return *this; // in place: the batch keeps its producer
    }
    template <class filter_f> MultiPipe &chain(Filter_GPU<filter_f> op)
    {
        attach(op); op.recycle_in = recycle_prev;
        auto rep = op.make_replica();
        auto *r = static_cast<typename Filter_GPU<filter_f>::Replica *>(rep.get());
        recycle_prev = [r](void *b) { r->take_back(b); };
        link(std::move(rep)); tail_is_gpu = true; return *this;
    }
    template <class map_f, class key_f> MultiPipe &chain(Map_GPU_KB<map_f, key_f> op)
    {
        attach(op); link(op.make_replica()); tail_is_gpu = true; return *this; // in place: the batch keeps its producer
    }
    template <class filter_f, class key_f> MultiPipe &chain(Filter_GPU_KB<filter_f, key_f> op)
    {
        attach(op); op.recycle_in = recycle_prev;
        auto rep = op.make_replica();
        auto *r = static_cast<typename Filter_GPU_KB<filter_f, key_f>::Replica *>(rep.get());
        recycle_prev = [r](void *b) { r->take_back(b); };
        link(std::move(rep)); tail_is_gpu = true; return *this;
    }
    template <class red_f, class key_f> MultiPipe &chain(Reduce_GPU<red_f, key_f> op)
    {
        attach(op); op.recycle_in = recycle_prev;
        auto rep = op.make_replica();
        auto *r = static_cast<typename Reduce_GPU<red_f, key_f>::Replica *>(rep.get());
        using T = typename Reduce_GPU<red_f, key_f>::tuple_t;
        recycle_prev = [r](void *b) { r->pool.put(reinterpret_cast<Batch_GPU_t<T> *>(b)); };
        link(std::move(rep)); tail_is_gpu = true; return *this;
    }
    template <class l_f, class c_f, class k_f> MultiPipe &chain(Ffat_Windows_GPU<l_f, c_f, k_f> op)
    {
        attach(op); op.recycle_in = recycle_prev;
        auto rep = op.make_replica();
        auto *r = static_cast<typename Ffat_Windows_GPU<l_f, c_f, k_f>::Replica *>(rep.get());
        using R = typename Ffat_Windows_GPU<l_f, c_f, k_f>::result_t;
        recycle_prev = [r](void *b) { r->pool.put(reinterpret_cast<Batch_GPU_t<R> *>(b)); };
        link(std::move(rep)); tail_is_gpu = true; return *this;
    }
    template <class sink_f> MultiPipe &chain_sink(Sink<sink_f> op)
    {
        if (has_sink) wf_fatal("MultiPipe is already terminated by a Sink");
        using T = std::decay_t<decltype(*std::declval<fn_arg_t<sink_f, 0>>())>; // std::optional<tuple_t> & -> tuple_t
        link(std::make_unique<SinkStage<T, sink_f>>(op.func, recycle_prev));
        op_names.push_back(op.getName()); has_sink = true; return *this;
    }
    template <class sink_f> MultiPipe &add_sink(Sink<sink_f> op) { return chain_sink(op); }
    size_t getNumStages() const { return stages.size(); }
};

class PipeGraph {
    std::string name; Execution_Mode_t mode; Time_Policy_t policy;
    std::vector<std::unique_ptr<MultiPipe>> pipes;
public:
    PipeGraph(std::string n, Execution_Mode_t m = Execution_Mode_t::DEFAULT, Time_Policy_t p = Time_Policy_t::INGRESS_TIME): name(std::move(n)), mode(m), policy(p)
    {
        if (wfb_device_count() <= 0) wf_fatal("no CUDA device: the GPU operators have no CPU fallback");
    }
    template <class src_f> MultiPipe &add_source(Source<src_f> src)
    {
        using T = typename shipper_tuple<std::decay_t<fn_arg_t<src_f, 0>>>::type;
        if (src.getOutputBatchSize() == 0) wf_fatal("Source feeding GPU operators must be built withOutputBatchSize(n > 0)"); // multipipe.hpp:486-488
        auto mp = std::make_unique<MultiPipe>();
        auto st = std::make_unique<SourceStage<T>>(std::function<void(Source_Shipper<T> &)>(src.func), src.getOutputBatchSize());
        SourceStage<T> *sp = st.get();
        mp->run_source = [sp]() { sp->run(); };
        mp->recycle_prev = [sp](void *b) { sp->recycle(reinterpret_cast<Batch_GPU_t<T> *>(b)); };
        mp->prevOutputBatchSize = src.getOutputBatchSize();
        mp->op_names.push_back(src.getName());
        mp->link(std::move(st));
        pipes.push_back(std::move(mp));
        return *pipes.back();
    }
    void run()
    {
        if (mode != Execution_Mode_t::DEFAULT) wf_fatal("GPU operators can only be used in DEFAULT mode");
        for (auto &p : pipes) { if (!p->has_sink) wf_fatal("MultiPipe without a Sink"); p->run_source(); }
        gpuErrChk(cudaDeviceSynchronize());
    }
    size_t getNumThreads() const { return 1; }
};

} // namespace wf
