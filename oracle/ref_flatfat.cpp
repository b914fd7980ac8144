/*
 * ref_flatfat.cpp -- thin C driver around the REFERENCE's own wf/flatfat.hpp (TEST INFRASTRUCTURE ONLY).
 *
 * Compiled by oracle/Makefile straight from /root/reference/wf (never copied into this repo) into
 * oracle/_ref/libwfref_flatfat.so. wf/flatfat.hpp is used unmodified; the only stand-in is the forward
 * declaration of wf::get_tuple_t_Comb that normally comes from wf/meta.hpp:472-494 (meta.hpp drags in
 * FastFlow, which is not vendored: /root/reference/CMakeLists.txt:48-54 git-clones it).
 *
 * What it pins: the oracle's restatement of FlatFAT insert / remove / getResult (oracle/wf_oracle.c,
 * fatcpu_*) and, through it, the window contents of the count-based FFAT operators. The per-tuple trigger
 * loop below restates FFAT_Replica::process_input_cb (wf/ffat_replica.hpp:215-278) and
 * eosnotifyCBWindows (:406-427) because ffat_replica.hpp itself needs FastFlow.
 */
#include <cstdint>
#include <cstdlib>
#include <type_traits>
#include <unordered_map>
#include <vector>
#include <basic.hpp>
#include <context.hpp>
namespace wf { // stands in for meta.hpp:472-494
template<typename F, typename A> A get_tuple_t_Comb(void (F::*)(const A &, const A &, A &) const);
template<typename F> decltype(get_tuple_t_Comb(&F::operator())) get_tuple_t_Comb(F);
}
#include <flatfat.hpp>

struct res_t
{
    uint64_t key; uint64_t id; int64_t isum; double fsum;
    res_t(): key(0), id(0), isum(0), fsum(0.0) {}
    res_t(uint64_t _key, uint64_t _id): key(_key), id(_id), isum(0), fsum(0.0) {}
};

struct Comb
{
    void operator()(const res_t &a, const res_t &b, res_t &out) const
    {
        int64_t is = a.isum + b.isum; double fs = a.fsum + b.fsum;
        out.isum = is; out.fsum = fs;
    }
};

using fat_t = wf::FlatFAT<Comb, uint64_t>;

struct KeyD
{
    fat_t fat;
    std::vector<res_t> pending;
    uint64_t rcv_counter = 0, slide_counter = 0, next_lwid = 0;
    KeyD(Comb *c, uint64_t key, uint64_t win, wf::RuntimeContext *ctx): fat(c, key, false, win, ctx) {}
};

struct RefFfatCpu
{
    Comb comb;
    wf::RuntimeContext ctx;
    uint64_t W, S, last_time = 0;
    std::unordered_map<uint64_t, KeyD> keyMap;
    RefFfatCpu(uint64_t w, uint64_t s): W(w), S(s) {}
};

extern "C" {

void *wfref_ffat_cpu_create(uint64_t win, uint64_t slide) { return new RefFfatCpu(win, slide); }
void wfref_ffat_cpu_destroy(void *h) { delete reinterpret_cast<RefFfatCpu *>(h); }

uint64_t wfref_ffat_cpu_process(void *hh, const res_t *res, uint64_t n, uint64_t wm,
                                res_t *out, uint64_t *out_ts, uint64_t out_cap)
{
    RefFfatCpu *h = reinterpret_cast<RefFfatCpu *>(hh);
    uint64_t nout = 0;
    h->last_time = wm;
    for (uint64_t i = 0; i < n; i++) {
        uint64_t key = res[i].key;
        auto it = h->keyMap.find(key);
        if (it == h->keyMap.end()) {
            it = h->keyMap.emplace(std::piecewise_construct, std::forward_as_tuple(key),
                                   std::forward_as_tuple(&h->comb, key, h->W, &h->ctx)).first;
        }
        KeyD &k = it->second;
        k.rcv_counter++; k.slide_counter++;
        res_t r(key, 0); r.isum = res[i].isum; r.fsum = res[i].fsum;
        k.pending.push_back(r);
        bool fired = false;
        if (k.rcv_counter == h->W) fired = true;
        else if (k.rcv_counter > h->W && (k.slide_counter % h->S == 0)) fired = true;
        if (fired) {
            uint64_t gwid = k.next_lwid++; k.slide_counter = 0;
            k.fat.insert(k.pending); k.pending.clear();
            res_t o = k.fat.getResult(gwid);
            k.fat.remove(h->S);
            if (nout < out_cap) { out[nout] = o; out_ts[nout] = wm; }
            nout++;
        }
    }
    return nout;
}

uint64_t wfref_ffat_cpu_eos(void *hh, res_t *out, uint64_t *out_ts, uint64_t out_cap)
{
    RefFfatCpu *h = reinterpret_cast<RefFfatCpu *>(hh);
    uint64_t nout = 0;
    for (auto &p : h->keyMap) {
        KeyD &k = p.second;
        k.fat.insert(k.pending); k.pending.clear();
        while (!k.fat.is_Empty()) {
            uint64_t gwid = k.next_lwid++;
            res_t o = k.fat.getResult(gwid);
            k.fat.remove(h->S);
            if (nout < out_cap) { out[nout] = o; out_ts[nout] = h->last_time; }
            nout++;
        }
    }
    return nout;
}

/* CPU baseline pipe over the REFERENCE FlatFAT: restated Map (wf/map.hpp:174-190) -> Filter (wf/filter.hpp:184-205)
 * loops in front of the reference wf::FlatFAT driven as FFAT_Replica::process_input_cb does, on one key shard
 * (key % nshards == shard, wf/keyby_emitter.hpp:215-217). Tuple layout = the 64-byte bench tuple of SURVEY 8d. */
struct tuple64_t { uint64_t key, id; int64_t ivalue; double fvalue; uint64_t pad[4]; };
struct RefPipe
{
    RefFfatCpu ffat;
    int map_kind, filt_kind; int64_t ia, im; double fa; uint32_t shard, nshards;
    std::vector<res_t> lift, out; std::vector<uint64_t> ots;
    RefPipe(uint64_t w, uint64_t s): ffat(w, s) {}
};

void *wfref_cpu_pipe_create(int map_kind, int64_t ia, double fa, int filt_kind, int64_t im,
                            uint64_t win, uint64_t slide, uint32_t shard, uint32_t nshards)
{
    RefPipe *p = new RefPipe(win, slide);
    p->map_kind = map_kind; p->ia = ia; p->fa = fa; p->filt_kind = filt_kind; p->im = im; p->shard = shard; p->nshards = nshards;
    return p;
}
void wfref_cpu_pipe_destroy(void *pp) { delete reinterpret_cast<RefPipe *>(pp); }

uint64_t wfref_ffat_cpu_process(void *hh, const res_t *res, uint64_t n, uint64_t wm, res_t *out, uint64_t *out_ts, uint64_t out_cap);

uint64_t wfref_cpu_pipe_run(void *pp, const tuple64_t *tuples, const uint64_t *ts, uint64_t n, uint64_t batch, int64_t *checksum)
{
    RefPipe *p = reinterpret_cast<RefPipe *>(pp);
    if (p->lift.size() < batch) { p->lift.resize(batch); p->out.resize(batch + 16); p->ots.resize(batch + 16); }
    uint64_t nwin = 0; int64_t cs = 0;
    for (uint64_t off = 0; off < n; off += batch) {
        uint64_t m = (n - off < batch) ? (n - off) : batch, nl = 0;
        for (uint64_t i = 0; i < m; i++) {
            const tuple64_t &src = tuples[off + i];
            if (src.key % p->nshards != p->shard) continue;
            tuple64_t t = src;
            if (p->map_kind == 1) { t.ivalue += p->ia; t.fvalue *= p->fa; }
            bool keep = p->filt_kind == 0 ? true : (p->filt_kind == 1 ? ((t.ivalue & 1) == 0) : ((t.ivalue % p->im) == 0));
            if (!keep) continue;
            res_t r(t.key, 0); r.isum = t.ivalue; r.fsum = t.fvalue;
            p->lift[nl++] = r;
        }
        uint64_t k = wfref_ffat_cpu_process(&p->ffat, p->lift.data(), nl, ts[off], p->out.data(), p->ots.data(), batch + 16);
        for (uint64_t i = 0; i < k && i < batch + 16; i++) cs += p->out[i].isum;
        nwin += k;
    }
    *checksum += cs;
    return nwin;
}

/* Raw FlatFAT access for unit-level pinning: one tree, scripted insert/remove/getResult. */
void *wfref_fat_create(uint64_t key, uint64_t n)
{
    static Comb comb; static wf::RuntimeContext ctx;
    return new fat_t(&comb, key, false, n, &ctx);
}
void wfref_fat_destroy(void *f) { delete reinterpret_cast<fat_t *>(f); }
void wfref_fat_insert(void *f, const res_t *in, uint64_t cnt)
{
    std::vector<res_t> v(in, in + cnt);
    reinterpret_cast<fat_t *>(f)->insert(v);
}
void wfref_fat_remove(void *f, uint64_t cnt) { reinterpret_cast<fat_t *>(f)->remove(cnt); }
void wfref_fat_result(void *f, uint64_t gwid, res_t *out) { *out = reinterpret_cast<fat_t *>(f)->getResult(gwid); }

} // extern "C"
