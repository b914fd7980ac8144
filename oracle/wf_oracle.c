/*
 * wf_oracle.c -- CPU restatement of the WindFlow GPU-operator hot path (TEST INFRASTRUCTURE ONLY).
 *
 * This file is the parity ORACLE of the repo: a plain-C restatement of the reference's algorithms for
 * Map_GPU / Filter_GPU / Reduce_GPU / KeyBy_Emitter_GPU grouping / Ffat_Windows_GPU (count-based and
 * time-based) and of the reference's CPU Ffat_Windows replica. Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it -- and only as the checker or the reported
 * CPU baseline, never as the product path (the product path is windflow_b200/libwfb200.so and fails
 * loudly when the CUDA library is missing).
 *
 * Parity pin: the window aggregation part is checked against the reference's own wf/flatfat.hpp compiled
 * unmodified (oracle/ref_flatfat.cpp -> oracle/_ref/) and, on the GPU box, against the reference's own
 * wf/flatfat_gpu.hpp compiled for sm_100a (oracle/ref_flatfat_gpu.cu -> oracle/_ref/). The reference holds
 * no golden vectors for this path; tests/golden/ holds vectors generated from these
 * reference builds by tests/golden/make_golden.py.
 *
 * All file:line citations are relative to /root/reference/wf/.
 *
 * Data model (schema independent): tuples are handled column-wise. A "lifted result" is
 *   wfo_res_t { key, id, isum, fsum }   (32 bytes)
 * which covers result32_t of the bench stream and the reference tests' result_t{key,id,value}
 * (tests/win_tests_gpu/win_common_gpu.hpp:61-80, value -> isum).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { uint64_t key; uint64_t id; int64_t ivalue; double fvalue; uint64_t pad[4]; } wfo_tuple64_t;
typedef struct { uint64_t key; uint64_t id; int64_t isum; double fsum; } wfo_res_t;

/* ------------------------------------------------------------------------------------------------
 * Synthetic stream. Identical integer arithmetic to the device generator in
 * windflow_b200/csrc/wfb_stream.cuh.
 * ---------------------------------------------------------------------------------------------- */
static inline uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

/* key_mode: 0 = i & (nkeys-1) style round robin (i % nkeys); 1 = splitmix64(i) % nkeys (uniform);
 *           2 = zipf via inverse-CDF table (cdf has nkeys entries in [0,1], ascending). */
void wfo_gen_tuple64(uint64_t seed, uint64_t start, uint64_t n, int key_mode, uint64_t nkeys,
                     const double *zipf_cdf, wfo_tuple64_t *out, uint64_t *ts)
{
    for (uint64_t j = 0; j < n; j++) {
        uint64_t i = start + j;
        wfo_tuple64_t t;
        memset(&t, 0, sizeof(t));
        if (key_mode == 0) t.key = i % nkeys;
        else if (key_mode == 1) t.key = splitmix64(i) % nkeys;
        else {
            double u = (double)(splitmix64(i ^ 0xA5A5A5A5A5A5A5A5ull) >> 11) * (1.0 / 9007199254740992.0);
            uint64_t lo = 0, hi = nkeys - 1; /* first index with cdf[idx] > u */
            while (lo < hi) { uint64_t mid = (lo + hi) >> 1; if (zipf_cdf[mid] > u) hi = mid; else lo = mid + 1; }
            t.key = lo;
        }
        t.id = i;
        t.ivalue = (int64_t)(splitmix64(seed ^ i) & 0xFFFFull);
        t.fvalue = (double)(splitmix64(seed ^ ~i) >> 11) * (1.0 / 9007199254740992.0);
        out[j] = t;
        ts[j] = i;
    }
}

/* Survivors of a SAMPLE of keys in [start, start+n) of the synthetic stream after the bench functors (map kind 1: ivalue += ia,
 * fvalue *= fa; filter kind 1: (ivalue & 1) == 0): for the check of the full-size bench configuration (bench.py --check), where
 * replaying 10^9 tuples through the window oracle is not an option but the history of a few keys is cheap to rebuild from the
 * generator. sel[k] != 0 marks a selected key (sel has nkeys_total entries). Outputs, in stream order: the key, the stream index,
 * the mapped ivalue / fvalue. Returns the number of survivors found (at most cap are written). */
uint64_t wfo_scan_keys(uint64_t seed, uint64_t start, uint64_t n, int key_mode, uint64_t nkeys, const uint8_t *sel, int64_t ia, double fa,
                       uint64_t *out_key, uint64_t *out_idx, int64_t *out_ival, double *out_fval, uint64_t cap)
{
    uint64_t m = 0;
    for (uint64_t j = 0; j < n; j++) {
        const uint64_t i = start + j;
        const uint64_t key = key_mode == 0 ? i % nkeys : splitmix64(i) % nkeys;
        if (!sel[key]) continue;
        const int64_t iv = (int64_t)(splitmix64(seed ^ i) & 0xFFFFull) + ia;
        if (iv & 1) continue;
        if (m < cap) {
            out_key[m] = key; out_idx[m] = i; out_ival[m] = iv;
            out_fval[m] = (double)(splitmix64(seed ^ ~i) >> 11) * (1.0 / 9007199254740992.0) * fa;
        }
        m++;
    }
    return m;
}

/* ------------------------------------------------------------------------------------------------
 * Map (map.hpp:174-190 in-place version; map_gpu.hpp:61-76 applies func to every item of the batch).
 * kind 0: identity; kind 1: ivalue += ia, fvalue *= fa (bench functor, and "+2" of
 * tests/graph_tests_gpu/graph_common_gpu.hpp:245-253 with fa = 1).
 * ---------------------------------------------------------------------------------------------- */
void wfo_map(int64_t *ival, double *fval, uint64_t n, int kind, int64_t ia, double fa)
{
    if (kind == 0) return;
    for (uint64_t i = 0; i < n; i++) { ival[i] += ia; if (fval) fval[i] *= fa; }
}

/* ------------------------------------------------------------------------------------------------
 * Filter predicate (filter.hpp:184-205: predicate true => keep; filter_gpu.hpp:72-88 writes flags[i]).
 * kind 0: keep all; kind 1: (ivalue & 1) == 0 (bench); kind 2: ivalue % im == 0
 * (tests/graph_tests_gpu/graph_common_gpu.hpp:198-215, C++ truncating %). Returns number kept.
 * Compaction itself is stable (thrust::copy_if, filter_gpu.hpp:555): survivors keep arrival order.
 * ---------------------------------------------------------------------------------------------- */
uint64_t wfo_filter_mask(const int64_t *ival, uint64_t n, int kind, int64_t im, uint8_t *mask)
{
    uint64_t kept = 0;
    for (uint64_t i = 0; i < n; i++) {
        uint8_t m;
        if (kind == 0) m = 1;
        else if (kind == 1) m = ((ival[i] & 1) == 0);
        else m = ((ival[i] % im) == 0);
        mask[i] = m;
        kept += m;
    }
    return kept;
}

/* ------------------------------------------------------------------------------------------------
 * Key grouping (keyby_emitter_gpu.hpp). map_idxs[i] = next index with the same key or -1;
 * start_idxs[k] = first index of the k-th distinct key; dist_keys[k] = that key.
 * order 0: first-seen order of distinct keys (CPU->GPU path, keyby_emitter_gpu.hpp:452-463);
 * order 1: ascending key order (GPU->GPU path, sort_by_key + unique_by_key_copy, :547-564).
 * Returns the number of distinct keys.
 * ---------------------------------------------------------------------------------------------- */
typedef struct { uint64_t key; uint32_t idx; } wfo_ki_t;
static int cmp_ki(const void *a, const void *b)
{
    const wfo_ki_t *x = (const wfo_ki_t *)a, *y = (const wfo_ki_t *)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    return x->idx < y->idx ? -1 : (x->idx > y->idx);
}

uint64_t wfo_keyby_group(const uint64_t *keys, uint64_t n, int order,
                         int32_t *start_idxs, int32_t *map_idxs, uint64_t *dist_keys)
{
    if (n == 0) return 0;
    wfo_ki_t *ki = (wfo_ki_t *)malloc(sizeof(wfo_ki_t) * n);
    for (uint64_t i = 0; i < n; i++) { ki[i].key = keys[i]; ki[i].idx = (uint32_t)i; }
    qsort(ki, n, sizeof(wfo_ki_t), cmp_ki); /* (key, idx) lexicographic == stable sort by key */
    uint64_t nk = 0;
    for (uint64_t i = 0; i < n; i++) {
        int last = (i == n - 1) || (ki[i].key != ki[i + 1].key);
        map_idxs[ki[i].idx] = last ? -1 : (int32_t)ki[i + 1].idx; /* Compute_Mapping_Kernel :84-100 */
        if (i == 0 || ki[i].key != ki[i - 1].key) {
            start_idxs[nk] = (int32_t)ki[i].idx; dist_keys[nk] = ki[i].key; nk++;
        }
    }
    if (order == 0) { /* re-order distinct keys by first occurrence */
        wfo_ki_t *fk = (wfo_ki_t *)malloc(sizeof(wfo_ki_t) * nk);
        for (uint64_t k = 0; k < nk; k++) { fk[k].key = (uint64_t)start_idxs[k]; fk[k].idx = (uint32_t)k; }
        qsort(fk, nk, sizeof(wfo_ki_t), cmp_ki);
        uint64_t *dk = (uint64_t *)malloc(sizeof(uint64_t) * nk);
        for (uint64_t k = 0; k < nk; k++) dk[k] = dist_keys[fk[k].idx];
        for (uint64_t k = 0; k < nk; k++) { start_idxs[k] = (int32_t)fk[k].key; dist_keys[k] = dk[k]; }
        free(dk); free(fk);
    }
    free(ki);
    return nk;
}

/* key -> destination (keyby_emitter.hpp:215-217, keyby_emitter_gpu.hpp:621): std::hash<size_t> is the
 * identity in libstdc++ (asserted by tests/win_tests/win_common.hpp:184), so dest = key % num_dests. */
void wfo_route(const uint64_t *keys, uint64_t n, uint32_t num_dests, uint32_t *dest)
{
    for (uint64_t i = 0; i < n; i++) dest[i] = (uint32_t)(keys[i] % num_dests);
}

/* ------------------------------------------------------------------------------------------------
 * Reduce_GPU, keyed, per batch (reduce_gpu.hpp:209-262): keys extracted, items sorted by key, then
 * reduce_by_key with func(lhs, rhs) and ts = max(lhs.ts, rhs.ts) (:88-105). Output: one item per
 * distinct key, ascending key order. A key seen once passes through untouched (no func call), so the
 * caller gets seg_first[k] (index of the first item of the segment in arrival order) and seg_len[k] to
 * rebuild non-aggregated fields. Fold order here is arrival order (left fold); thrust's association is
 * implementation-defined, irrelevant for integers, tolerance for floating point.
 * Returns the number of distinct keys.
 * ---------------------------------------------------------------------------------------------- */
uint64_t wfo_reduce_by_key(const uint64_t *keys, const int64_t *ival, const double *fval, const uint64_t *ts,
                           uint64_t n, uint64_t *out_keys, int64_t *out_isum, double *out_fsum,
                           uint64_t *out_ts, uint32_t *seg_first, uint32_t *seg_len)
{
    if (n == 0) return 0;
    wfo_ki_t *ki = (wfo_ki_t *)malloc(sizeof(wfo_ki_t) * n);
    for (uint64_t i = 0; i < n; i++) { ki[i].key = keys[i]; ki[i].idx = (uint32_t)i; }
    qsort(ki, n, sizeof(wfo_ki_t), cmp_ki);
    uint64_t nk = 0;
    for (uint64_t i = 0; i < n; ) {
        uint64_t j = i;
        int64_t is = ival[ki[i].idx];
        double fs = fval ? fval[ki[i].idx] : 0.0;
        uint64_t mts = ts[ki[i].idx];
        for (j = i + 1; j < n && ki[j].key == ki[i].key; j++) {
            is += ival[ki[j].idx];
            if (fval) fs += fval[ki[j].idx];
            if (mts < ts[ki[j].idx]) mts = ts[ki[j].idx];
        }
        out_keys[nk] = ki[i].key; out_isum[nk] = is; out_fsum[nk] = fs; out_ts[nk] = mts;
        seg_first[nk] = ki[i].idx; seg_len[nk] = (uint32_t)(j - i);
        nk++;
        i = j;
    }
    free(ki);
    return nk;
}

/* ------------------------------------------------------------------------------------------------
 * Combine (associative). tests: output.value = in1.value + in2.value
 * (tests/win_tests_gpu/win_common_gpu.hpp:306-314); bench: field-wise + on isum/fsum. key/id of the
 * output are left untouched, exactly like the reference functor.
 * ---------------------------------------------------------------------------------------------- */
static inline void comb(const wfo_res_t *a, const wfo_res_t *b, wfo_res_t *out)
{
    int64_t is = a->isum + b->isum; double fs = a->fsum + b->fsum; /* tolerate out aliasing a or b */
    out->isum = is; out->fsum = fs;
}
static inline wfo_res_t res_init(uint64_t key, uint64_t id)
{
    wfo_res_t r; r.key = key; r.id = id; r.isum = 0; r.fsum = 0.0; return r; /* result_t(key, id) */
}

/* ------------------------------------------------------------------------------------------------
 * FlatFAT_GPU restated on the CPU (flatfat_gpu.hpp:142-424). One instance per key.
 * Deviation (documented in DESIGN.md): leaves [B, n) are initialised to result_t() instead of being
 * left as uninitialised device memory (flatfat_gpu.hpp:184-191 never writes them, :343-356 reads them).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    uint64_t B, Nb, W, S, n, treeSize;
    uint64_t offset, incr;
    wfo_res_t *tree;
} fatgpu_t;

static inline uint64_t parent_pos(uint64_t pos, uint64_t n) { return (pos >> 1) | n; } /* flatfat_gpu.hpp:55-58 */

static void fatgpu_init(fatgpu_t *f, uint64_t B, uint64_t Nb, uint64_t W, uint64_t S)
{
    uint64_t n = 1; while (n < B) n <<= 1; /* flatfat_gpu.hpp:184-185 */
    f->B = B; f->Nb = Nb; f->W = W; f->S = S; f->n = n; f->treeSize = 2 * n - 1;
    f->offset = 0; f->incr = 0;
    f->tree = (wfo_res_t *)calloc(f->treeSize, sizeof(wfo_res_t));
}

static void fatgpu_add_cb(fatgpu_t *f, const wfo_res_t *data, uint64_t size) /* flatfat_gpu.hpp:226-252 */
{
    uint64_t pos = (f->offset + f->incr) % f->B;
    uint64_t spaceLeft = f->B - pos;
    if (size <= spaceLeft) memcpy(f->tree + pos, data, size * sizeof(wfo_res_t));
    else {
        memcpy(f->tree + pos, data, spaceLeft * sizeof(wfo_res_t));
        memcpy(f->tree, data + spaceLeft, (size - spaceLeft) * sizeof(wfo_res_t));
    }
    f->incr += size;
}

static void fatgpu_build(fatgpu_t *f) /* flatfat_gpu.hpp:338-359, Init_TreeLevel_Kernel :62-72 */
{
    wfo_res_t *A = f->tree;
    uint64_t pw = 1;
    wfo_res_t *Bl = A + f->n / pw;
    uint64_t i = f->n / 2;
    while (Bl < f->tree + f->treeSize && i > 0) {
        for (uint64_t k = 0; k < i; k++) comb(&A[2 * k], &A[2 * k + 1], &Bl[k]);
        A = Bl; pw <<= 1; Bl = A + f->n / pw; i /= 2;
    }
    f->incr = 0;
}

static void fatgpu_update(fatgpu_t *f, uint64_t num_new) /* flatfat_gpu.hpp:362-395, Update_TreeLevel_Kernel :76-89 */
{
    uint64_t pw = 1;
    wfo_res_t *A = f->tree;
    wfo_res_t *Bl = A + f->n / pw;
    uint64_t sizeB = (f->B + (pw << 1) - 1) / (pw << 1);
    uint64_t update_pos = parent_pos(f->offset, f->n);
    uint64_t numSeen = f->n;
    uint64_t distance = update_pos - numSeen;
    uint64_t sizeUpdate = (num_new + (pw << 1) - 1) / (pw << 1) + 1;
    while (Bl < f->tree + f->treeSize) {
        for (uint64_t i = 0; i < sizeUpdate; i++) {
            uint64_t my_i = (i + distance) % sizeB;
            comb(&A[my_i * 2], &A[my_i * 2 + 1], &Bl[my_i]);
        }
        pw <<= 1; A = Bl; Bl = A + f->n / pw;
        sizeB = (f->B + (pw << 1) - 1) / (pw << 1);
        update_pos = parent_pos(update_pos, f->n);
        numSeen += f->n / pw;
        distance = update_pos - numSeen;
        sizeUpdate = (num_new + (pw << 1) - 1) / (pw << 1) + 1;
    }
    f->offset = (f->offset + num_new) % f->B;
    f->incr = 0;
}

/* Compute_Results_Kernel, flatfat_gpu.hpp:93-139 */
static void fatgpu_results(const fatgpu_t *f, uint64_t key, uint64_t gwid0, uint64_t wm,
                           wfo_res_t *out, uint64_t *out_ts)
{
    int64_t B = (int64_t)f->B;
    for (uint64_t i = 0; i < f->Nb; i++) {
        int64_t wS = (int64_t)((f->offset + i * f->S) % f->B);
        int64_t WIN = (int64_t)f->W;
        out[i] = res_init(key, gwid0 + i);
        out_ts[i] = wm;
        while (WIN > 0) {
            int64_t range;
            wS = wS >= B ? 0 : wS;
            range = wS == 0 ? B : (wS & -wS);
            int64_t pw = WIN;
            pw |= pw >> 1; pw |= pw >> 2; pw |= pw >> 4; pw |= pw >> 8; pw |= pw >> 16; pw |= pw >> 32;
            pw = (pw >> 1) + 1;
            range = range < pw ? range : pw;
            int64_t tr = range; uint64_t tn = (uint64_t)wS;
            while (tr > 1) { tn = parent_pos(tn, f->n); tr >>= 1; }
            comb(&out[i], &f->tree[tn], &out[i]);
            int64_t oldWS = wS;
            wS += range;
            range = wS >= B ? B - oldWS : range;
            WIN -= range;
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * Ffat_Replica_GPU, count-based windows (ffat_replica_gpu.hpp:734-867).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    uint64_t key; int used;
    fatgpu_t fat;
    uint64_t next_gwid, count, count_triggerer; /* ffat_replica_gpu.hpp:445-448,461-466 */
    /* semantic shadow: every lifted result of the key, for the linear-fold cross check */
    wfo_res_t *hist; uint64_t hist_len, hist_cap;
} keydesc_t;

typedef struct {
    uint64_t W, S, Nb, B;
    int keep_history;
    keydesc_t *tab; uint64_t cap, used; /* open-addressing map key -> descriptor (keyMap, :514) */
} wfo_ffat_gpu_t;

static keydesc_t *kd_find(wfo_ffat_gpu_t *h, uint64_t key)
{
    if ((h->used + 1) * 2 > h->cap) { /* grow */
        uint64_t ncap = h->cap ? h->cap * 2 : 1024;
        keydesc_t *nt = (keydesc_t *)calloc(ncap, sizeof(keydesc_t));
        for (uint64_t i = 0; i < h->cap; i++) if (h->tab[i].used) {
            uint64_t p = splitmix64(h->tab[i].key) & (ncap - 1);
            while (nt[p].used) p = (p + 1) & (ncap - 1);
            nt[p] = h->tab[i];
        }
        free(h->tab); h->tab = nt; h->cap = ncap;
    }
    uint64_t p = splitmix64(key) & (h->cap - 1);
    while (h->tab[p].used && h->tab[p].key != key) p = (p + 1) & (h->cap - 1);
    if (!h->tab[p].used) {
        keydesc_t *k = &h->tab[p];
        memset(k, 0, sizeof(*k));
        k->used = 1; k->key = key;
        fatgpu_init(&k->fat, h->B, h->Nb, h->W, h->S);
        k->next_gwid = 0; k->count = 0; k->count_triggerer = h->B;
        h->used++;
    }
    return &h->tab[p];
}

wfo_ffat_gpu_t *wfo_ffat_gpu_create(uint64_t win, uint64_t slide, uint64_t nb, int keep_history)
{
    wfo_ffat_gpu_t *h = (wfo_ffat_gpu_t *)calloc(1, sizeof(*h));
    h->W = win; h->S = slide; h->Nb = nb; h->B = (nb - 1) * slide + win; /* ffat_replica_gpu.hpp:657 */
    h->keep_history = keep_history;
    return h;
}

void wfo_ffat_gpu_destroy(wfo_ffat_gpu_t *h)
{
    for (uint64_t i = 0; i < h->cap; i++) if (h->tab[i].used) { free(h->tab[i].fat.tree); free(h->tab[i].hist); }
    free(h->tab); free(h);
}

/* process_wins_cb, ffat_replica_gpu.hpp:830-867. Appends Nb results per trigger to out. */
static uint64_t process_wins_cb(wfo_ffat_gpu_t *h, keydesc_t *k, const wfo_res_t *res, uint64_t num,
                                uint64_t wm, wfo_res_t *out, uint64_t *out_ts, uint64_t out_cap, uint64_t nout)
{
    uint64_t off = 0;
    while (k->count + num >= k->count_triggerer) {
        uint64_t take = k->count_triggerer - k->count;
        fatgpu_add_cb(&k->fat, res + off, take);
        num -= take; off += take; k->count += take;
        if (k->count_triggerer == h->B) fatgpu_build(&k->fat);           /* first window, :838-849 */
        else fatgpu_update(&k->fat, h->S * h->Nb);                       /* :850-861 */
        if (nout + h->Nb <= out_cap) fatgpu_results(&k->fat, k->key, k->next_gwid, wm, out + nout, out_ts + nout);
        nout += h->Nb;
        k->next_gwid += h->Nb;
        k->count_triggerer += h->S * h->Nb;
    }
    if (num > 0) { fatgpu_add_cb(&k->fat, res + off, num); k->count += num; }
    return nout;
}

/* process_batch_cb (keyed), ffat_replica_gpu.hpp:734-801. Input: the lifted results of one batch in
 * arrival order (res[i].key = key of tuple i; Lifting_Kernel_CB_Keyed :108-121). Groups by key with a stable
 * ascending sort (:751) and walks the distinct keys in that order (:782-800). Returns the number of window
 * results this batch produces (written to out if it fits in out_cap). */
uint64_t wfo_ffat_gpu_process_batch(wfo_ffat_gpu_t *h, const wfo_res_t *res, uint64_t n, uint64_t wm,
                                    wfo_res_t *out, uint64_t *out_ts, uint64_t out_cap)
{
    if (n == 0) return 0;
    wfo_ki_t *ki = (wfo_ki_t *)malloc(sizeof(wfo_ki_t) * n);
    wfo_res_t *sorted = (wfo_res_t *)malloc(sizeof(wfo_res_t) * n);
    for (uint64_t i = 0; i < n; i++) { ki[i].key = res[i].key; ki[i].idx = (uint32_t)i; }
    qsort(ki, n, sizeof(wfo_ki_t), cmp_ki);
    for (uint64_t i = 0; i < n; i++) sorted[i] = res[ki[i].idx];
    uint64_t nout = 0;
    for (uint64_t i = 0; i < n; ) {
        uint64_t j = i + 1;
        while (j < n && ki[j].key == ki[i].key) j++;
        keydesc_t *k = kd_find(h, ki[i].key);
        if (h->keep_history) {
            if (k->hist_len + (j - i) > k->hist_cap) {
                k->hist_cap = (k->hist_len + (j - i)) * 2;
                k->hist = (wfo_res_t *)realloc(k->hist, k->hist_cap * sizeof(wfo_res_t));
            }
            memcpy(k->hist + k->hist_len, sorted + i, (j - i) * sizeof(wfo_res_t));
            k->hist_len += (j - i);
        }
        nout = process_wins_cb(h, k, sorted + i, j - i, wm, out, out_ts, out_cap, nout);
        i = j;
    }
    free(sorted); free(ki);
    return nout;
}

/* ------------------------------------------------------------------------------------------------
 * Ffat_Replica_GPU, time-based windows (ffat_replica_gpu.hpp:870-1047), PendingPanes_Queue (:263-420),
 * Aggregate_Panes_Kernel (:214-260), Lifting_Kernel_TB_Keyed (:150-171).
 * The queue is restated as a ring indexed by pane_id % capacity (the reference keeps first_pos / last_pos over the
 * same content). Defined where the reference only asserts: a key that has fewer pending panes than a group needs
 * (ffat_replica_gpu.hpp:1031,:1037 `assert`) gets empty panes (result_t()) for the missing ones.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    uint64_t key; int used;
    fatgpu_t fat;
    uint64_t q_first_id, q_num, q_cap; wfo_res_t *q_buf;        /* PendingPanes_Queue */
    uint64_t pane_id_triggerer, next_gwid; int firstWinDone;    /* Key_Descriptor :441-447, :462-464 */
} tbkey_t;

typedef struct {
    uint64_t win_p, slide_p, pane_len, lateness, Nb, Bp;
    tbkey_t *tab; uint64_t cap, used;
    uint64_t ignored;                                           /* ignored_tuples_gpu */
} wfo_ffat_tb_t;

static uint64_t gcd_u64(uint64_t a, uint64_t b) { while (b) { uint64_t t = a % b; a = b; b = t; } return a; }

wfo_ffat_tb_t *wfo_ffat_tb_create(uint64_t win, uint64_t slide, uint64_t lateness, uint64_t nb)
{
    wfo_ffat_tb_t *h = (wfo_ffat_tb_t *)calloc(1, sizeof(*h));
    h->pane_len = gcd_u64(win, slide);                          /* :639-642 */
    h->win_p = win / h->pane_len; h->slide_p = slide / h->pane_len;
    h->lateness = lateness; h->Nb = nb;
    h->Bp = (nb - 1) * h->slide_p + h->win_p;                   /* batchSize :657 */
    return h;
}

void wfo_ffat_tb_destroy(wfo_ffat_tb_t *h)
{
    for (uint64_t i = 0; i < h->cap; i++) if (h->tab[i].used) { free(h->tab[i].fat.tree); free(h->tab[i].q_buf); }
    free(h->tab); free(h);
}

uint64_t wfo_ffat_tb_ignored(const wfo_ffat_tb_t *h) { return h->ignored; }

static tbkey_t *tb_find(wfo_ffat_tb_t *h, uint64_t key)
{
    if ((h->used + 1) * 2 > h->cap) {
        uint64_t ncap = h->cap ? h->cap * 2 : 1024;
        tbkey_t *nt = (tbkey_t *)calloc(ncap, sizeof(tbkey_t));
        for (uint64_t i = 0; i < h->cap; i++) if (h->tab[i].used) {
            uint64_t p = splitmix64(h->tab[i].key) & (ncap - 1);
            while (nt[p].used) p = (p + 1) & (ncap - 1);
            nt[p] = h->tab[i];
        }
        free(h->tab); h->tab = nt; h->cap = ncap;
    }
    uint64_t p = splitmix64(key) & (h->cap - 1);
    while (h->tab[p].used && h->tab[p].key != key) p = (p + 1) & (h->cap - 1);
    if (!h->tab[p].used) {
        tbkey_t *k = &h->tab[p];
        memset(k, 0, sizeof(*k));
        k->used = 1; k->key = key;
        fatgpu_init(&k->fat, h->Bp, h->Nb, h->win_p, h->slide_p);
        k->q_cap = h->Bp; k->q_buf = (wfo_res_t *)calloc(k->q_cap, sizeof(wfo_res_t)); /* initial capacity :470 */
        k->pane_id_triggerer = h->Bp - 1;                       /* :463 */
        h->used++;
    }
    return &h->tab[p];
}

static void tbq_resize(tbkey_t *k, uint64_t ncap) /* PendingPanes_Queue::resize :326-357 (content preserved) */
{
    wfo_res_t *nb = (wfo_res_t *)calloc(ncap, sizeof(wfo_res_t));
    for (uint64_t j = 0; j < k->q_num; j++) nb[(k->q_first_id + j) % ncap] = k->q_buf[(k->q_first_id + j) % k->q_cap];
    free(k->q_buf); k->q_buf = nb; k->q_cap = ncap;
}

typedef struct { uint64_t key, pane; uint32_t idx; } wfo_kpi_t;
static int cmp_kpi(const void *a, const void *b) /* lessThan_func_gpu_t :174-192: key ascending, pane DEscending; stable (idx) */
{
    const wfo_kpi_t *x = (const wfo_kpi_t *)a, *y = (const wfo_kpi_t *)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    if (x->pane != y->pane) return x->pane > y->pane ? -1 : 1;
    return x->idx < y->idx ? -1 : (x->idx > y->idx ? 1 : 0);
}

/* One input batch: res[i] = lift(tuple i) with res[i].key = key of tuple i, ts[i] its timestamp, wm the batch watermark.
 * Appends Nb results per fired group (ts of a result = wm). Returns the number of results. */
uint64_t wfo_ffat_tb_process_batch(wfo_ffat_tb_t *h, const wfo_res_t *res, const uint64_t *ts, uint64_t n, uint64_t wm,
                                   wfo_res_t *out, uint64_t *out_ts, uint64_t out_cap)
{
    if (n == 0) return 0;
    const uint64_t F = wm >= h->lateness ? (wm - h->lateness) / h->pane_len : 0; /* first_pane_not_complete :875-881 */
    wfo_kpi_t *ki = (wfo_kpi_t *)malloc(sizeof(wfo_kpi_t) * n);
    for (uint64_t i = 0; i < n; i++) {
        ki[i].key = res[i].key; ki[i].pane = ts[i] / h->pane_len; ki[i].idx = (uint32_t)i;
        if (ki[i].pane < F) h->ignored++;                        /* :165-167 (statistic only) */
    }
    qsort(ki, n, sizeof(wfo_kpi_t), cmp_kpi);                   /* sort_by_key :917-921 */
    /* reduce_by_key over equal (key, pane), arrival order inside (:925-935) */
    wfo_res_t *pv = (wfo_res_t *)malloc(sizeof(wfo_res_t) * n);
    uint64_t *pp = (uint64_t *)malloc(sizeof(uint64_t) * n), *pk = (uint64_t *)malloc(sizeof(uint64_t) * n);
    uint64_t np = 0;
    for (uint64_t i = 0; i < n; ) {
        uint64_t j = i + 1;
        wfo_res_t acc = res[ki[i].idx];
        while (j < n && ki[j].key == ki[i].key && ki[j].pane == ki[i].pane) { comb(&acc, &res[ki[j].idx], &acc); j++; } /* thrust_comb_func_t: comb(acc, next, acc) */
        pv[np] = acc; pp[np] = ki[i].pane; pk[np] = ki[i].key; np++;
        i = j;
    }
    uint64_t nout = 0;
    for (uint64_t i = 0; i < np; ) { /* one key at a time, ascending (:963-987) */
        uint64_t j = i + 1;
        while (j < np && pk[j] == pk[i]) j++;
        tbkey_t *k = tb_find(h, pk[i]);
        /* push_panes(new_panes, new_infos, num, newest_pane_id = pp[i]) :360-391 */
        const uint64_t newest = pp[i];
        if (newest >= k->q_first_id) {
            const uint64_t need = newest - k->q_first_id + 2;
            if (need > k->q_cap) tbq_resize(k, need);
        }
        for (uint64_t a = i; a < j; a++) { /* Aggregate_Panes_Kernel :214-260 */
            const uint64_t pane = pp[a];
            if (pane < k->q_first_id) continue; /* late pane */
            wfo_res_t *slot = &k->q_buf[pane % k->q_cap];
            if (k->q_num > 0 && pane < k->q_first_id + k->q_num) comb(slot, &pv[a], slot); /* :236 */
            else {
                *slot = pv[a];
                const uint64_t lower = (a == j - 1) ? (k->q_first_id + k->q_num) : (pp[a + 1] + 1); /* first missing pane below */
                for (uint64_t m = (lower > k->q_first_id + k->q_num ? lower : k->q_first_id + k->q_num); m < pane; m++)
                    memset(&k->q_buf[m % k->q_cap], 0, sizeof(wfo_res_t)); /* result_t() */
            }
        }
        if (newest >= k->q_first_id + k->q_num) k->q_num = newest - k->q_first_id + 1;
        /* process_wins_tb :1022-1047 with next_pane_id = F (:985) */
        while (k->pane_id_triggerer < F) {
            const uint64_t need = k->firstWinDone ? h->slide_p * h->Nb : h->Bp;
            wfo_res_t *grp = (wfo_res_t *)calloc(need, sizeof(wfo_res_t));
            for (uint64_t m = 0; m < need; m++) /* pop_and_add :394-415; missing panes are empty */
                if (m < k->q_num) grp[m] = k->q_buf[(k->q_first_id + m) % k->q_cap];
            fatgpu_add_cb(&k->fat, grp, need);
            free(grp);
            k->q_first_id += need; k->q_num = k->q_num > need ? k->q_num - need : 0;
            if (!k->firstWinDone) { fatgpu_build(&k->fat); k->firstWinDone = 1; }
            else fatgpu_update(&k->fat, h->slide_p * h->Nb);
            if (nout + h->Nb <= out_cap) fatgpu_results(&k->fat, k->key, k->next_gwid, wm, out + nout, out_ts + nout);
            nout += h->Nb;
            k->next_gwid += h->Nb;
            k->pane_id_triggerer += h->slide_p * h->Nb;
        }
        i = j;
    }
    free(pk); free(pp); free(pv); free(ki);
    return nout;
}

/* Semantic definition of window `gwid` of `key`: left fold of comb over the key's lifted
 * results [gwid*S, gwid*S+W) starting from result_t(key, gwid). Needs keep_history=1. Returns 0 if the
 * window is not complete yet. */
int wfo_ffat_gpu_window_linear(wfo_ffat_gpu_t *h, uint64_t key, uint64_t gwid, wfo_res_t *out)
{
    keydesc_t *k = kd_find(h, key);
    uint64_t a = gwid * h->S, b = a + h->W;
    if (!h->keep_history || k->hist_len < b) return 0;
    wfo_res_t r = res_init(key, gwid);
    for (uint64_t i = a; i < b; i++) comb(&r, &k->hist[i], &r);
    *out = r;
    return 1;
}

/* ------------------------------------------------------------------------------------------------
 * CPU Ffat_Windows replica, count-based (ffat_replica.hpp:215-278 process_input_cb, :406-427 EOS) over
 * the CPU FlatFAT (flatfat.hpp:53-348) restated. This is BASELINE.json config 1's window stage and the
 * CPU baseline the bench reports. One instance == one replica (single thread).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    uint64_t n, front, back; int isEmpty;
    wfo_res_t *tree; /* 2n entries, root at 1, leaves at [n, 2n) */
    uint64_t key;
} fatcpu_t;

static void fatcpu_init(fatcpu_t *f, uint64_t win, uint64_t key) /* flatfat.hpp:159-177 */
{
    uint64_t n = 1; while (n < win) n <<= 1;
    f->n = n; f->front = n - 1; f->back = n - 1; f->isEmpty = 1; f->key = key;
    f->tree = (wfo_res_t *)malloc(sizeof(wfo_res_t) * 2 * n);
    for (uint64_t i = 0; i < 2 * n; i++) f->tree[i] = res_init(key, 0);
}

/* shared tail of insert(vector)/remove(count): bottom-up refresh of a FIFO list of dirty parents
 * (flatfat.hpp:224-241, :283-300). The list is a ring of capacity 2n (each node enters at most once per call). */
typedef struct { uint64_t *q; uint64_t head, tail, cap; } nodeq_t;
static inline void nq_push(nodeq_t *q, uint64_t v) { q->q[q->tail % q->cap] = v; q->tail++; }
static inline int nq_empty(const nodeq_t *q) { return q->head == q->tail; }
static inline uint64_t nq_back(const nodeq_t *q) { return q->q[(q->tail - 1) % q->cap]; }

static void fatcpu_refresh(fatcpu_t *f, nodeq_t *q)
{
    while (!nq_empty(q)) {
        uint64_t node = q->q[q->head % q->cap]; q->head++;
        wfo_res_t r = res_init(f->key, 0);
        comb(&f->tree[node << 1], &f->tree[(node << 1) + 1], &r);
        f->tree[node] = r;
        uint64_t p = node >> 1;
        if (node != 1 && (nq_empty(q) || nq_back(q) != p)) nq_push(q, p);
    }
}

static void fatcpu_insert(fatcpu_t *f, const wfo_res_t *in, uint64_t cnt, nodeq_t *q) /* flatfat.hpp:200-242 */
{
    q->head = q->tail = 0;
    for (uint64_t i = 0; i < cnt; i++) {
        if (f->front == f->back && f->front == f->n - 1) { f->front++; f->back++; f->isEmpty = 0; }
        else if (f->back == 2 * f->n - 1) { if (f->front != f->n) f->back = f->n; else abort(); }
        else if (f->front != f->back + 1) f->back++;
        else abort();
        f->tree[f->back] = in[i];
        uint64_t p = f->back >> 1;
        if (f->back != 1 && (nq_empty(q) || nq_back(q) != p)) nq_push(q, p);
    }
    fatcpu_refresh(f, q);
}

static void fatcpu_remove(fatcpu_t *f, uint64_t count, nodeq_t *q) /* flatfat.hpp:264-301 */
{
    q->head = q->tail = 0;
    for (uint64_t i = 0; i < count; i++) {
        f->tree[f->front] = res_init(f->key, 0);
        uint64_t p = f->front >> 1;
        if (f->front != 1 && (nq_empty(q) || nq_back(q) != p)) nq_push(q, p);
        if (f->front == f->back) { f->front = f->back = f->n - 1; f->isEmpty = 1; break; }
        else if (f->front == 2 * f->n - 1) f->front = f->n;
        else f->front++;
    }
    fatcpu_refresh(f, q);
}

static wfo_res_t fatcpu_prefix(const fatcpu_t *f, uint64_t pos) /* flatfat.hpp:83-107 */
{
    uint64_t i = pos; wfo_res_t acc = f->tree[pos];
    while (i != 1) {
        uint64_t p = i >> 1;
        if (i == (p << 1) + 1) { wfo_res_t tmp = acc; acc = res_init(f->key, 0); comb(&f->tree[p << 1], &tmp, &acc); }
        i = p;
    }
    return acc;
}
static wfo_res_t fatcpu_suffix(const fatcpu_t *f, uint64_t pos) /* flatfat.hpp:110-134 */
{
    uint64_t i = pos; wfo_res_t acc = f->tree[pos];
    while (i != 1) {
        uint64_t p = i >> 1;
        if (i == (p << 1)) { wfo_res_t tmp = acc; acc = res_init(f->key, 0); comb(&tmp, &f->tree[(p << 1) + 1], &acc); }
        i = p;
    }
    return acc;
}
static wfo_res_t fatcpu_result(const fatcpu_t *f, uint64_t gwid) /* flatfat.hpp:304-339, isCommutative=false */
{
    wfo_res_t res = res_init(f->key, gwid);
    if (f->front <= f->back) comb(&f->tree[1], &res, &res);
    else { wfo_res_t pr = fatcpu_prefix(f, f->back), sf = fatcpu_suffix(f, f->front); comb(&sf, &pr, &res); }
    return res;
}

typedef struct {
    uint64_t key; int used;
    fatcpu_t fat;
    wfo_res_t *pending; uint64_t npending, cap_pending;
    uint64_t rcv_counter, slide_counter, next_lwid;
} keydesc_cpu_t;

typedef struct {
    uint64_t W, S;
    keydesc_cpu_t *tab; uint64_t cap, used;
    nodeq_t q;
    uint64_t last_time;
} wfo_ffat_cpu_t;

wfo_ffat_cpu_t *wfo_ffat_cpu_create(uint64_t win, uint64_t slide)
{
    wfo_ffat_cpu_t *h = (wfo_ffat_cpu_t *)calloc(1, sizeof(*h));
    h->W = win; h->S = slide;
    uint64_t n = 1; while (n < win) n <<= 1;
    h->q.cap = 4 * n + 8; h->q.q = (uint64_t *)malloc(sizeof(uint64_t) * h->q.cap);
    return h;
}
void wfo_ffat_cpu_destroy(wfo_ffat_cpu_t *h)
{
    for (uint64_t i = 0; i < h->cap; i++) if (h->tab[i].used) { free(h->tab[i].fat.tree); free(h->tab[i].pending); }
    free(h->tab); free(h->q.q); free(h);
}
static keydesc_cpu_t *kdc_find(wfo_ffat_cpu_t *h, uint64_t key)
{
    if ((h->used + 1) * 2 > h->cap) {
        uint64_t ncap = h->cap ? h->cap * 2 : 1024;
        keydesc_cpu_t *nt = (keydesc_cpu_t *)calloc(ncap, sizeof(keydesc_cpu_t));
        for (uint64_t i = 0; i < h->cap; i++) if (h->tab[i].used) {
            uint64_t p = splitmix64(h->tab[i].key) & (ncap - 1);
            while (nt[p].used) p = (p + 1) & (ncap - 1);
            nt[p] = h->tab[i];
        }
        free(h->tab); h->tab = nt; h->cap = ncap;
    }
    uint64_t p = splitmix64(key) & (h->cap - 1);
    while (h->tab[p].used && h->tab[p].key != key) p = (p + 1) & (h->cap - 1);
    if (!h->tab[p].used) {
        keydesc_cpu_t *k = &h->tab[p];
        memset(k, 0, sizeof(*k));
        k->used = 1; k->key = key;
        fatcpu_init(&k->fat, h->W, key);
        h->used++;
    }
    return &h->tab[p];
}

/* process_input_cb for n tuples that already went through lift (res[i].key = key). Emits complete windows
 * (ts = watermark, DEFAULT mode :268-269). Returns the new number of outputs. */
uint64_t wfo_ffat_cpu_process(wfo_ffat_cpu_t *h, const wfo_res_t *res, uint64_t n, uint64_t wm,
                              wfo_res_t *out, uint64_t *out_ts, uint64_t out_cap)
{
    uint64_t nout = 0;
    h->last_time = wm;
    for (uint64_t i = 0; i < n; i++) {
        keydesc_cpu_t *k = kdc_find(h, res[i].key);
        k->rcv_counter++; k->slide_counter++;
        if (k->npending == k->cap_pending) {
            k->cap_pending = k->cap_pending ? k->cap_pending * 2 : 64;
            k->pending = (wfo_res_t *)realloc(k->pending, sizeof(wfo_res_t) * k->cap_pending);
        }
        wfo_res_t r = res_init(res[i].key, 0); r.isum = res[i].isum; r.fsum = res[i].fsum; /* lift into result_t(key) */
        k->pending[k->npending++] = r;
        int fired = 0;
        if (k->rcv_counter == h->W) fired = 1;
        else if (k->rcv_counter > h->W && (k->slide_counter % h->S == 0)) fired = 1;
        if (fired) {
            uint64_t gwid = k->next_lwid++; k->slide_counter = 0;
            fatcpu_insert(&k->fat, k->pending, k->npending, &h->q); k->npending = 0;
            wfo_res_t o = fatcpu_result(&k->fat, gwid);
            fatcpu_remove(&k->fat, h->S, &h->q);
            if (nout < out_cap) { out[nout] = o; out_ts[nout] = wm; }
            nout++;
        }
    }
    return nout;
}

/* eosnotifyCBWindows, ffat_replica.hpp:406-427: flush partial windows (CPU operator only; the GPU operator
 * emits nothing at end of stream, ffat_replica_gpu.hpp:1050-1056). Keys are visited in table order (the
 * reference's order is unordered_map iteration order -- unspecified), callers sort by (key, gwid). */
uint64_t wfo_ffat_cpu_eos(wfo_ffat_cpu_t *h, wfo_res_t *out, uint64_t *out_ts, uint64_t out_cap)
{
    uint64_t nout = 0;
    for (uint64_t i = 0; i < h->cap; i++) if (h->tab[i].used) {
        keydesc_cpu_t *k = &h->tab[i];
        fatcpu_insert(&k->fat, k->pending, k->npending, &h->q); k->npending = 0;
        while (!k->fat.isEmpty) {
            uint64_t gwid = k->next_lwid++;
            wfo_res_t o = fatcpu_result(&k->fat, gwid);
            fatcpu_remove(&k->fat, h->S, &h->q);
            if (nout < out_cap) { out[nout] = o; out_ts[nout] = h->last_time; }
            nout++;
        }
    }
    return nout;
}

/* ------------------------------------------------------------------------------------------------
 * CPU baseline driver: the reference's CPU Map -> Filter -> Ffat_Windows(CB) path (map.hpp:174-190,
 * filter.hpp:184-205, ffat_replica.hpp:215-278) over pre-generated tuple64 buffers, on ONE key shard
 * (BASELINE.json config 1 shape). `shard`/`nshards` select keys with key % nshards == shard (keyby routing,
 * keyby_emitter.hpp:215-217), so `nshards` threads each owning one pipe mirror parallelism = nshards.
 * The watermark of each batch of `batch` tuples is its first timestamp. State persists across
 * wfo_cpu_pipe_run calls (a stream fed buffer by buffer). Returns the number of windows of the call;
 * *checksum accumulates isum of every window.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    wfo_ffat_cpu_t *ffat;
    int map_kind, filt_kind; int64_t ia, im; double fa;
    uint32_t shard, nshards;
    wfo_res_t *lift, *out; uint64_t *ots; uint64_t cap;
} wfo_cpu_pipe_t;

wfo_cpu_pipe_t *wfo_cpu_pipe_create(int map_kind, int64_t ia, double fa, int filt_kind, int64_t im,
                                    uint64_t win, uint64_t slide, uint32_t shard, uint32_t nshards)
{
    wfo_cpu_pipe_t *p = (wfo_cpu_pipe_t *)calloc(1, sizeof(*p));
    p->ffat = wfo_ffat_cpu_create(win, slide);
    p->map_kind = map_kind; p->ia = ia; p->fa = fa; p->filt_kind = filt_kind; p->im = im;
    p->shard = shard; p->nshards = nshards;
    return p;
}

void wfo_cpu_pipe_destroy(wfo_cpu_pipe_t *p)
{
    wfo_ffat_cpu_destroy(p->ffat); free(p->lift); free(p->out); free(p->ots); free(p);
}

uint64_t wfo_cpu_pipe_run(wfo_cpu_pipe_t *p, const wfo_tuple64_t *tuples, const uint64_t *ts, uint64_t n,
                          uint64_t batch, int64_t *checksum)
{
    if (p->cap < batch) {
        p->cap = batch;
        p->lift = (wfo_res_t *)realloc(p->lift, sizeof(wfo_res_t) * batch);
        p->out = (wfo_res_t *)realloc(p->out, sizeof(wfo_res_t) * (batch + 16));
        p->ots = (uint64_t *)realloc(p->ots, sizeof(uint64_t) * (batch + 16));
    }
    uint64_t nwin = 0; int64_t cs = 0;
    for (uint64_t off = 0; off < n; off += batch) {
        uint64_t m = (n - off < batch) ? (n - off) : batch;
        uint64_t nl = 0;
        for (uint64_t i = 0; i < m; i++) {
            const wfo_tuple64_t *src = &tuples[off + i];
            if (src->key % p->nshards != p->shard) continue;
            wfo_tuple64_t t = *src;
            if (p->map_kind == 1) { t.ivalue += p->ia; t.fvalue *= p->fa; }
            int keep = p->filt_kind == 0 ? 1 : (p->filt_kind == 1 ? ((t.ivalue & 1) == 0) : ((t.ivalue % p->im) == 0));
            if (!keep) continue;
            p->lift[nl].key = t.key; p->lift[nl].id = 0; p->lift[nl].isum = t.ivalue; p->lift[nl].fsum = t.fvalue; nl++;
        }
        uint64_t k = wfo_ffat_cpu_process(p->ffat, p->lift, nl, ts[off], p->out, p->ots, batch + 16);
        for (uint64_t i = 0; i < k && i < batch + 16; i++) cs += p->out[i].isum;
        nwin += k;
    }
    *checksum += cs;
    return nwin;
}

#ifdef __cplusplus
}
#endif
