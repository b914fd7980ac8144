// ff/ff.hpp -- a small, header-only, FastFlow-API-compatible thread/queue runtime (C++17, std::thread).
//
// WindFlow's host side is written over FastFlow (ff_node / ff_monode / ff_minode / ff_pipeline / ff_a2a, combine_with_*,
// MPMC_Ptr_Queue): wf/basic_operator.hpp:54, wf/multipipe.hpp:96-660, wf/basic_emitter.hpp:65. FastFlow itself is a separate
// project that the reference fetches at configure time (CMakeLists.txt:48-54) and is not available to this repository, so the
// builder API of include/wf/windflow_gpu.hpp runs over this runtime instead. It implements, from scratch, exactly the subset
// of the FastFlow programming model WindFlow's graphs use:
//   - a node is an object with svc_init / svc / eosnotify / svc_end; svc returns GO_ON, EOS or a task to forward;
//   - ff_pipeline chains stages, ff_a2a connects every node of a first set to every node of a second set, both nest;
//   - combine_with_firststage / combine_with_laststage fuse two nodes into one thread (the consumer's svc is called directly
//     from the producer's ff_send_out: "chaining");
//   - every edge of the flattened graph is a bounded single-producer/single-consumer ring of pointers; a node with several
//     inputs polls them round-robin and sees the channel of the current task through get_channel_id();
//   - end of stream: a producer that terminates sends one EOS mark on each of its output channels; a consumer calls
//     eosnotify(channel) for each, and terminates when all its inputs have delivered theirs.
// One extension is used by the GPU replicas of this repository: ff_node::ff_poll() lets svc() take further tasks that are
// already queued on the node's input (a replica that finds K batches waiting hands all of them to one launch sequence).
//
// The same header lets the unmodified reference (wf/windflow.hpp + wf/windflow_gpu.hpp) compile and run for the parity
// tests (oracle/Makefile, -I include): that use is test infrastructure; nothing here comes from FastFlow's sources.
#pragma once
#include <atomic>
#include <cassert>
#include <chrono>
#include <climits>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <memory>
#include <thread>
#include <vector>
#include <sys/types.h>
#include <sys/time.h>
#include <unistd.h> // (programs written over FastFlow get getopt & co. through its headers)

#ifndef DEFAULT_BUFFER_CAPACITY
#define DEFAULT_BUFFER_CAPACITY 2048
#endif

namespace ff {

class ff_node;
class ff_pipeline;
class ff_a2a;
class ff_group;

namespace rt {

// bounded single-producer / single-consumer ring of pointers (one per graph edge)
class Channel {
    static constexpr size_t CAP = 1024; // tasks in flight per edge; a full ring makes the producer wait (back-pressure)
    alignas(64) std::atomic<size_t> head{0};
    alignas(64) std::atomic<size_t> tail{0};
    alignas(64) void *slots[CAP];
public:
    bool try_push(void *p)
    {
        const size_t t = tail.load(std::memory_order_relaxed);
        if (t - head.load(std::memory_order_acquire) == CAP) return false;
        slots[t % CAP] = p;
        tail.store(t + 1, std::memory_order_release);
        return true;
    }
    bool peek(void **p)
    {
        const size_t h = head.load(std::memory_order_relaxed);
        if (h == tail.load(std::memory_order_acquire)) return false;
        *p = slots[h % CAP];
        return true;
    }
    void pop() { head.store(head.load(std::memory_order_relaxed) + 1, std::memory_order_release); }
};

inline void backoff(unsigned &spins)
{
    if (++spins < 64) return;
    if (spins < 256) { std::this_thread::yield(); return; }
    std::this_thread::sleep_for(std::chrono::microseconds(spins < 2048 ? 20 : 200));
}

struct Worker; // one thread: a chain of combined nodes, its input and output channels

} // namespace rt

// ---- nodes -----------------------------------------------------------------------------------------------------------------
class ff_node {
    friend struct rt::Worker;
    friend class ff_pipeline;
    friend class ff_a2a;
    friend class ff_comb;
    friend class ff_group;
protected:
    enum kind_t { LEAF, COMB, PIPE, A2A, GROUP };
    rt::Worker *worker_ = nullptr; // set when the graph is flattened
    size_t pos_ = 0;               // position in the worker's chain
    bool skipfirstpop_ = false;
    virtual kind_t kind() const { return LEAF; }
public:
    static inline void *const EOS = reinterpret_cast<void *>(ULLONG_MAX);
    static inline void *const GO_ON = reinterpret_cast<void *>(ULLONG_MAX - 1);
    static inline void *const GO_OUT = reinterpret_cast<void *>(ULLONG_MAX - 2);
    virtual ~ff_node() {}
    virtual int svc_init() { return 0; }
    virtual void *svc(void *task) = 0;
    virtual void svc_end() {}
    virtual void eosnotify(ssize_t /*id*/ = -1) {}
    void skipfirstpop(bool sk) { skipfirstpop_ = sk; }
    // forward a task: to the next node of the chain when this node is combined with a successor, else to an output channel
    // (round-robin over the channels when there are several)
    inline bool ff_send_out(void *task, int id = -1, unsigned long retry = ((unsigned long) -1), unsigned long ticks = 0);
    // the next task already queued on this node's input, if any (never an end-of-stream mark). Only from inside svc().
    inline bool ff_poll(void **task);
    inline ssize_t get_my_id() const;
    virtual bool isMultiInput() const { return false; }
    virtual bool isMultiOutput() const { return false; }
    double ffTime() const { return 0.0; }
    virtual void ffStats(std::ostream &) {}
};

class ff_monode: public ff_node { // multi-output node
public:
    inline bool ff_send_out_to(void *task, int id, unsigned long retry = ((unsigned long) -1), unsigned long ticks = 0);
    inline size_t get_num_outchannels() const;
    bool isMultiOutput() const override { return true; }
};

class ff_minode: public ff_node { // multi-input node
public:
    inline ssize_t get_channel_id() const;
    inline size_t get_num_inchannels() const;
    bool isMultiInput() const override { return true; }
};

// two (or more) nodes fused into one thread: the output of node i is the input of node i+1 (combine_with_*)
class ff_comb: public ff_node {
    friend struct rt::Worker;
    friend class ff_pipeline;
    std::vector<ff_node *> nodes;
    std::vector<ff_node *> owned;
protected:
    kind_t kind() const override { return COMB; }
public:
    ~ff_comb() override { for (auto *n : owned) delete n; }
    void *svc(void *) override { return EOS; } // never called: the worker runs the chain
    void append(ff_node *n, bool cleanup)
    {
        if (auto *c = dynamic_cast<ff_comb *>(n)) { for (auto *x : c->nodes) nodes.push_back(x); for (auto *x : c->owned) owned.push_back(x); c->nodes.clear(); c->owned.clear(); if (cleanup) owned.push_back(c); }
        else { nodes.push_back(n); if (cleanup) owned.push_back(n); }
    }
    const std::vector<ff_node *> &chain() const { return nodes; }
};

namespace rt {

struct Graph;

struct Worker {
    std::vector<ff_node *> chain;
    std::vector<Channel *> in, out;
    std::thread th;
    ssize_t cur_channel = -1;
    size_t rr_out = 0, rr_in = 0;
    bool stop = false; // a node of the chain returned EOS from svc
    size_t id = 0;
    std::vector<bool> closed; // inputs that delivered their end-of-stream mark

    void push(Channel *c, void *task)
    {
        unsigned spins = 0;
        while (!c->try_push(task)) backoff(spins);
    }
    void emit(void *task, int ch)
    {
        if (out.empty()) return; // (a last stage that still emits: nothing listens)
        if (ch < 0) { ch = static_cast<int>(rr_out); rr_out = (rr_out + 1) % out.size(); }
        push(out[static_cast<size_t>(ch) % out.size()], task);
    }
    // result of node `pos` (or a task it sends): on to node pos + 1, or out of the worker
    void forward(size_t pos, void *task, int ch)
    {
        if (pos + 1 < chain.size()) feed(pos + 1, task);
        else emit(task, ch);
    }
    void feed(size_t pos, void *task)
    {
        void *r = chain[pos]->svc(task);
        if (r == ff_node::GO_ON || r == ff_node::GO_OUT) return;
        if (r == ff_node::EOS) { stop = true; return; }
        forward(pos, r, -1);
    }
    // next data task queued on the inputs (never consumes an end-of-stream mark)
    bool poll(void **task)
    {
        for (size_t k = 0; k < in.size(); k++) {
            const size_t c = (rr_in + k) % in.size();
            void *p;
            if (closed[c] || !in[c]->peek(&p) || p == ff_node::EOS) continue;
            in[c]->pop();
            cur_channel = static_cast<ssize_t>(c); rr_in = (c + 1) % in.size();
            *task = p;
            return true;
        }
        return false;
    }
    void run()
    {
        for (size_t i = 0; i < chain.size(); i++) {
            chain[i]->worker_ = this; chain[i]->pos_ = i;
        }
        for (auto *n : chain) if (n->svc_init() < 0) { std::fprintf(stderr, "ff: svc_init failed\n"); std::abort(); }
        closed.assign(in.size(), false);
        size_t nclosed = 0;
        if (in.empty() || chain[0]->skipfirstpop_) { // a source (or a node asking to run before its first input)
            do {
                void *r = chain[0]->svc(nullptr);
                if (r == ff_node::EOS) { stop = true; break; }
                if (r != ff_node::GO_ON && r != ff_node::GO_OUT) forward(0, r, -1);
            } while (in.empty() && !stop);
        }
        unsigned spins = 0;
        while (!stop && nclosed < in.size()) {
            bool got = false;
            for (size_t k = 0; k < in.size() && !stop; k++) {
                const size_t c = (rr_in + k) % in.size();
                void *p;
                if (closed[c] || !in[c]->peek(&p)) continue;
                in[c]->pop();
                got = true;
                if (p == ff_node::EOS) { closed[c] = true; nclosed++; chain[0]->eosnotify(static_cast<ssize_t>(c)); continue; }
                cur_channel = static_cast<ssize_t>(c); rr_in = (c + 1) % in.size();
                feed(0, p);
                break; // fairness: restart the scan after the channel just served
            }
            if (got) spins = 0; else backoff(spins);
        }
        // end of stream: the first node has seen eosnotify(channel) for every input; the nodes combined after it flush now, in order
        for (size_t i = 1; i < chain.size(); i++) chain[i]->eosnotify(-1);
        for (auto *c : out) push(c, ff_node::EOS);
        for (auto *n : chain) n->svc_end();
    }
};

struct Ends { std::vector<Worker *> ins, outs; };

struct Graph {
    std::vector<std::unique_ptr<Worker>> workers;
    std::vector<std::unique_ptr<Channel>> channels;
    Worker *add_worker(const std::vector<ff_node *> &chain)
    {
        workers.emplace_back(new Worker());
        workers.back()->chain = chain; workers.back()->id = workers.size() - 1;
        return workers.back().get();
    }
    void connect(const std::vector<Worker *> &outs, const std::vector<Worker *> &ins)
    {
        for (auto *o : outs) for (auto *i : ins) {
            channels.emplace_back(new Channel());
            o->out.push_back(channels.back().get()); i->in.push_back(channels.back().get());
        }
    }
    void start() { for (auto &w : workers) w->th = std::thread([p = w.get()] { p->run(); }); }
    void join() { for (auto &w : workers) if (w->th.joinable()) w->th.join(); }
};

} // namespace rt

inline bool ff_node::ff_send_out(void *task, int id, unsigned long, unsigned long)
{
    assert(worker_ != nullptr);
    worker_->forward(pos_, task, id);
    return true;
}
inline bool ff_node::ff_poll(void **task) { return worker_ != nullptr && pos_ == 0 && worker_->poll(task); }
inline ssize_t ff_node::get_my_id() const { return worker_ ? static_cast<ssize_t>(worker_->id) : -1; }
inline bool ff_monode::ff_send_out_to(void *task, int id, unsigned long, unsigned long)
{
    assert(worker_ != nullptr);
    worker_->forward(pos_, task, id);
    return true;
}
inline size_t ff_monode::get_num_outchannels() const { return worker_ ? worker_->out.size() : 0; }
inline ssize_t ff_minode::get_channel_id() const { return worker_ ? worker_->cur_channel : -1; }
inline size_t ff_minode::get_num_inchannels() const { return worker_ ? worker_->in.size() : 0; }

// ---- containers ----------------------------------------------------------------------------------------------------------------
class ff_a2a: public ff_node {
    friend class ff_pipeline;
    std::vector<ff_node *> first, second;
    std::vector<ff_node *> cleanup_list;
protected:
    kind_t kind() const override { return A2A; }
public:
    explicit ff_a2a(bool = false, int = 0, int = 0, bool = false) {}
    ~ff_a2a() override { for (auto *n : cleanup_list) delete n; }
    void *svc(void *) override { return EOS; }
    int add_firstset(const std::vector<ff_node *> &w, int /*ondemand*/ = 0, bool cleanup = false)
    {
        first = w;
        if (cleanup) for (auto *n : w) cleanup_list.push_back(n);
        return 0;
    }
    int add_secondset(const std::vector<ff_node *> &w, bool cleanup = false)
    {
        second = w;
        if (cleanup) for (auto *n : w) cleanup_list.push_back(n);
        return 0;
    }
    int change_secondset(const std::vector<ff_node *> &w, bool cleanup = false, bool remove_from_cleanup = false)
    {
        if (remove_from_cleanup) remove_from_cleanuplist(second);
        second = w;
        if (cleanup) for (auto *n : w) cleanup_list.push_back(n);
        return 0;
    }
    void remove_from_cleanuplist(const std::vector<ff_node *> &w)
    {
        for (auto *n : w) for (size_t i = 0; i < cleanup_list.size(); i++) if (cleanup_list[i] == n) { cleanup_list.erase(cleanup_list.begin() + i); break; }
    }
    const std::vector<ff_node *> &getFirstSet() const { return first; }
    const std::vector<ff_node *> &getSecondSet() const { return second; }
    bool isMultiInput() const override { return true; }
    bool isMultiOutput() const override { return true; }
};

// Extension (not in FastFlow): the replicas of one operator as a pipeline stage -- every member is an input and an output of
// the stage, so consecutive groups of a pipeline are connected all-to-all (what a matrioska of ff_a2a's expresses by nesting).
class ff_group: public ff_node {
    friend class ff_pipeline;
    std::vector<ff_node *> members;
    std::vector<ff_node *> cleanup_list;
protected:
    kind_t kind() const override { return GROUP; }
public:
    ~ff_group() override { for (auto *n : cleanup_list) delete n; }
    void *svc(void *) override { return EOS; }
    void add(ff_node *n, bool cleanup = false) { members.push_back(n); if (cleanup) cleanup_list.push_back(n); }
    const std::vector<ff_node *> &getMembers() const { return members; }
};

class ff_pipeline: public ff_node {
    friend class ff_a2a;
    template <class T> friend int combine_with_firststage(ff_pipeline &, T *, bool);
    template <class T> friend int combine_with_laststage(ff_pipeline &, T *, bool);
    std::vector<ff_node *> stages;
    std::vector<ff_node *> cleanup_list;
    std::unique_ptr<rt::Graph> graph; // set by run() on the outermost pipeline
protected:
    kind_t kind() const override { return PIPE; }
    static rt::Ends build(ff_node *n, rt::Graph &g)
    {
        rt::Ends e;
        switch (n->kind()) {
        case LEAF: { rt::Worker *w = g.add_worker({n}); e.ins = {w}; e.outs = {w}; break; }
        case COMB: { rt::Worker *w = g.add_worker(static_cast<ff_comb *>(n)->chain()); e.ins = {w}; e.outs = {w}; break; }
        case PIPE: {
            auto *p = static_cast<ff_pipeline *>(n);
            bool firstStage = true;
            for (auto *s : p->stages) {
                rt::Ends se = build(s, g);
                if (firstStage) { e = se; firstStage = false; }
                else { g.connect(e.outs, se.ins); e.outs = se.outs; }
            }
            break;
        }
        case GROUP: {
            for (auto *s : static_cast<ff_group *>(n)->getMembers()) { rt::Ends se = build(s, g); e.ins.insert(e.ins.end(), se.ins.begin(), se.ins.end()); e.outs.insert(e.outs.end(), se.outs.begin(), se.outs.end()); }
            break;
        }
        case A2A: {
            auto *a = static_cast<ff_a2a *>(n);
            rt::Ends l, r;
            for (auto *s : a->first) { rt::Ends se = build(s, g); l.ins.insert(l.ins.end(), se.ins.begin(), se.ins.end()); l.outs.insert(l.outs.end(), se.outs.begin(), se.outs.end()); }
            for (auto *s : a->second) { rt::Ends se = build(s, g); r.ins.insert(r.ins.end(), se.ins.begin(), se.ins.end()); r.outs.insert(r.outs.end(), se.outs.begin(), se.outs.end()); }
            g.connect(l.outs, r.ins);
            e.ins = l.ins; e.outs = r.outs;
            break;
        }
        }
        return e;
    }
public:
    explicit ff_pipeline(bool = false, int = 0, int = 0, bool = false) {}
    ~ff_pipeline() override { for (auto *n : cleanup_list) delete n; }
    void *svc(void *) override { return EOS; }
    int add_stage(ff_node *s, bool cleanup = false)
    {
        stages.push_back(s);
        if (cleanup) cleanup_list.push_back(s);
        return 0;
    }
    int remove_stage(int pos)
    {
        if (pos < 0 || static_cast<size_t>(pos) >= stages.size()) return -1;
        stages.erase(stages.begin() + pos); // (a removed stage stays on the cleanup list, as the reference expects: wf/multipipe.hpp:527)
        return 0;
    }
    const std::vector<ff_node *> &getStages() const { return stages; }
    // number of threads the flattened graph runs on (every leaf or combined node is one)
    int cardinality() const
    {
        rt::Graph g;
        build(const_cast<ff_pipeline *>(this), g);
        return static_cast<int>(g.workers.size());
    }
    int run(bool = false)
    {
        graph.reset(new rt::Graph());
        build(this, *graph);
        graph->start();
        return 0;
    }
    int wait() { if (graph) graph->join(); return 0; }
    int wait_freezing() { return wait(); }
    int run_and_wait_end() { if (run() < 0) return -1; return wait(); }
    int run_then_freeze() { return run(); }
    bool isMultiInput() const override { return !stages.empty() && stages.front()->isMultiInput(); }
    bool isMultiOutput() const override { return !stages.empty() && stages.back()->isMultiOutput(); }
};

// node -> first stage of the pipeline, fused in one thread
template <class T>
int combine_with_firststage(ff_pipeline &pipe, T *node, bool cleanup)
{
    if (pipe.stages.empty()) return -1;
    auto *c = new ff_comb();
    c->append(node, cleanup);
    c->append(pipe.stages.front(), false);
    pipe.stages.front() = c;
    pipe.cleanup_list.push_back(c);
    return 0;
}
// last stage of the pipeline -> node, fused in one thread
template <class T>
int combine_with_laststage(ff_pipeline &pipe, T *node, bool cleanup)
{
    if (pipe.stages.empty()) return -1;
    auto *c = new ff_comb();
    c->append(pipe.stages.back(), false);
    c->append(node, cleanup);
    pipe.stages.back() = c;
    pipe.cleanup_list.push_back(c);
    return 0;
}

// ---- bounded multi-producer / multi-consumer queue of pointers (batch recycling, wf/recycling.hpp) ---------------------------
class MPMC_Ptr_Queue {
    struct Cell { std::atomic<size_t> seq; void *data; };
    std::unique_ptr<Cell[]> buf;
    size_t mask = 0;
    alignas(64) std::atomic<size_t> enq{0};
    alignas(64) std::atomic<size_t> deq{0};
public:
    bool init(size_t size)
    {
        size_t n = 2;
        while (n < size) n <<= 1;
        buf.reset(new Cell[n]);
        for (size_t i = 0; i < n; i++) buf[i].seq.store(i, std::memory_order_relaxed);
        mask = n - 1;
        return true;
    }
    bool push(void *const data)
    {
        size_t pos = enq.load(std::memory_order_relaxed);
        for (;;) {
            Cell &c = buf[pos & mask];
            const size_t seq = c.seq.load(std::memory_order_acquire);
            const intptr_t dif = static_cast<intptr_t>(seq) - static_cast<intptr_t>(pos);
            if (dif == 0) { if (enq.compare_exchange_weak(pos, pos + 1, std::memory_order_relaxed)) { c.data = data; c.seq.store(pos + 1, std::memory_order_release); return true; } }
            else if (dif < 0) return false; // full
            else pos = enq.load(std::memory_order_relaxed);
        }
    }
    bool pop(void **data)
    {
        size_t pos = deq.load(std::memory_order_relaxed);
        for (;;) {
            Cell &c = buf[pos & mask];
            const size_t seq = c.seq.load(std::memory_order_acquire);
            const intptr_t dif = static_cast<intptr_t>(seq) - static_cast<intptr_t>(pos + 1);
            if (dif == 0) { if (deq.compare_exchange_weak(pos, pos + 1, std::memory_order_relaxed)) { *data = c.data; c.seq.store(pos + mask + 1, std::memory_order_release); return true; } }
            else if (dif < 0) return false; // empty
            else pos = deq.load(std::memory_order_relaxed);
        }
    }
};

} // namespace ff
