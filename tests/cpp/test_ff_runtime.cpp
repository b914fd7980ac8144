// Unit test of include/ff/ff.hpp (the FastFlow-API-compatible runtime): nested pipeline / all-to-all graphs, combined nodes,
// channel ids, end-of-stream notification order, round-robin and addressed sends, ff_poll, the MPMC queue.
// Build: g++ -std=c++17 -O2 -pthread -I include tests/cpp/test_ff_runtime.cpp -o tests/cpp/test_ff_runtime.bin
#include <ff/ff.hpp>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <set>
#include <vector>

#define CHECK(c) do { if (!(c)) { std::fprintf(stderr, "CHECK failed: %s (line %d)\n", #c, __LINE__); std::exit(1); } } while (0)

struct Item { long v; int src; };

struct Source: ff::ff_monode {
    int id; long n;
    Source(int i, long n_): id(i), n(n_) {}
    void *svc(void *) override
    {
        for (long i = 0; i < n; i++) {
            Item *it = new Item{i, id};
            ff_send_out_to(it, static_cast<int>(i % get_num_outchannels())); // addressed send: item i -> destination i % n
        }
        return EOS;
    }
};

// multi-input first half of a combined worker: records the channel of every task and the end-of-stream notifications
struct Collector: ff::ff_minode {
    std::vector<long> per_channel; std::vector<int> eos_ids; int inits = 0;
    int svc_init() override { per_channel.assign(get_num_inchannels(), 0); inits++; return 0; }
    void *svc(void *t) override { per_channel[static_cast<size_t>(get_channel_id())]++; return t; } // a returned task goes to the next node
    void eosnotify(ssize_t id) override { eos_ids.push_back(static_cast<int>(id)); }
};

struct Worker: ff::ff_monode {
    long seen = 0; int eos_calls = 0; bool flushed_before_end = false, ended = false;
    std::vector<Item *> held;
    void *svc(void *t) override
    {
        seen++;
        held.push_back(static_cast<Item *>(t)); // keeps two tasks back, like an emitter with an open batch
        if (held.size() > 2) { ff_send_out(held.front()); held.erase(held.begin()); }
        return GO_ON;
    }
    void eosnotify(ssize_t) override { eos_calls++; for (auto *i : held) ff_send_out(i); held.clear(); flushed_before_end = !ended; }
    void svc_end() override { ended = true; }
};

struct Sink: ff::ff_minode {
    long sum = 0, count = 0, polled = 0; std::set<int> sources;
    void *svc(void *t) override
    {
        void *more = t;
        do { // drain what is already queued (the extension the GPU replicas use)
            Item *i = static_cast<Item *>(more);
            sum += i->v; count++; sources.insert(i->src);
            delete i;
            if (more != t) polled++;
        } while (ff_poll(&more));
        return GO_ON;
    }
};

struct Killer: ff::ff_minode {
    int svc_init() override { skipfirstpop(true); return 0; }
    void *svc(void *) override { return EOS; }
};

int main()
{
    const int NS = 3, NW = 4; const long N = 20000;
    // pipeline{ a2a{ first: NS x pipe(source) ; second: pipe{ a2a{ first: NW x pipe(comb(collector, worker)) ; second: pipe{ a2a{ first: pipe(sink), second: pipe(killer) } } } } } }
    // -- the nesting wf/multipipe.hpp builds ("matrioska")
    std::vector<Source *> sources; std::vector<Collector *> colls; std::vector<Worker *> workers;
    Sink *sink = new Sink();
    ff::ff_pipeline top;
    auto *m1 = new ff::ff_a2a();
    std::vector<ff::ff_node *> fs1;
    for (int i = 0; i < NS; i++) { auto *p = new ff::ff_pipeline(); sources.push_back(new Source(i, N)); p->add_stage(sources.back(), true); fs1.push_back(p); }
    m1->add_firstset(fs1, 0, true);
    auto *m2 = new ff::ff_a2a();
    std::vector<ff::ff_node *> fs2;
    for (int i = 0; i < NW; i++) {
        auto *p = new ff::ff_pipeline();
        workers.push_back(new Worker()); colls.push_back(new Collector());
        p->add_stage(workers.back(), true);
        ff::combine_with_firststage(*p, colls.back(), true);
        fs2.push_back(p);
    }
    m2->add_firstset(fs2, 0, true);
    auto *m3 = new ff::ff_a2a();
    { auto *p = new ff::ff_pipeline(); p->add_stage(sink, true); m3->add_firstset({p}, 0, true); }
    { auto *p = new ff::ff_pipeline(); p->add_stage(new Killer(), true); m3->add_secondset({p}, true); }
    { auto *p = new ff::ff_pipeline(); p->add_stage(m3, true); m2->add_secondset({p}, true); }
    { auto *p = new ff::ff_pipeline(); p->add_stage(m2, true); m1->add_secondset({p}, true); }
    top.add_stage(m1, true);
    CHECK(top.cardinality() == NS + NW + 2);
    CHECK(top.run_and_wait_end() == 0);
    // every item arrived exactly once
    CHECK(sink->count == NS * N);
    CHECK(sink->sum == NS * (N * (N - 1) / 2));
    CHECK(sink->sources.size() == static_cast<size_t>(NS));
    for (int w = 0; w < NW; w++) {
        CHECK(colls[w]->inits == 1);
        CHECK(colls[w]->per_channel.size() == static_cast<size_t>(NS));
        for (int s = 0; s < NS; s++) CHECK(colls[w]->per_channel[s] == N / NW);   // addressed sends, channel = index of the source
        CHECK(colls[w]->eos_ids.size() == static_cast<size_t>(NS));                  // one notification per input channel
        std::set<int> ids(colls[w]->eos_ids.begin(), colls[w]->eos_ids.end());
        CHECK(ids.size() == static_cast<size_t>(NS));
        CHECK(workers[w]->seen == NS * N / NW);
        CHECK(workers[w]->eos_calls == 1 && workers[w]->flushed_before_end);          // the combined node flushes once, before svc_end
    }
    std::printf("graph ok: %ld items, %ld taken by ff_poll\n", sink->count, sink->polled);

    // combine_with_laststage: source and a filter in one thread
    struct Evens: ff::ff_node { void *svc(void *t) override { Item *i = static_cast<Item *>(t); if (i->v & 1) { delete i; return GO_ON; } return t; } };
    ff::ff_pipeline p2;
    auto *st = new ff::ff_pipeline();
    st->add_stage(new Source(0, 1000), true);
    ff::combine_with_laststage(*st, new Evens(), true);
    Sink *s2 = new Sink();
    p2.add_stage(st, true); p2.add_stage(s2, true);
    CHECK(p2.cardinality() == 2);
    CHECK(p2.run_and_wait_end() == 0);
    CHECK(s2->count == 500 && s2->sum == 2 * (499 * 500 / 2));

    // MPMC queue
    ff::MPMC_Ptr_Queue q; q.init(8);
    long a[8]; void *out = nullptr;
    CHECK(!q.pop(&out));
    for (int i = 0; i < 8; i++) CHECK(q.push(&a[i]));
    CHECK(!q.push(&a[0]));
    for (int i = 0; i < 8; i++) { CHECK(q.pop(&out)); CHECK(out == &a[i]); }
    CHECK(!q.pop(&out));
    std::printf("ff runtime OK\n");
    return 0;
}
