/*
 * ref_flatfat_gpu.cu -- thin C driver around the REFERENCE's own wf/flatfat_gpu.hpp (TEST INFRASTRUCTURE
 * ONLY; also "the reference's kernel on this box" for BASELINE.md row B-gpu-ref-fat).
 *
 * Compiled by oracle/Makefile straight from /root/reference/wf (never copied into this repo) with
 * nvcc -arch=sm_100a into oracle/_ref/libwfref_flatfat_gpu.so; it needs a GPU to RUN, so it is only called
 * from `-m gpu` tests and bench.py on the GPU box. wf/flatfat_gpu.hpp (FlatFAT_GPU + Init_TreeLevel_Kernel /
 * Update_TreeLevel_Kernel / Compute_Results_Kernel, :62-139) is used unmodified.
 *
 * The per-key trigger loop restates Ffat_Replica_GPU::process_wins_cb (wf/ffat_replica_gpu.hpp:830-867)
 * because ffat_replica_gpu.hpp itself needs FastFlow + TBB, which are not vendored.
 */
#include <cstdint>
#include <cstdio>
#include <iostream>
#include <cassert>
#include <cuda_runtime.h>
#include <basic.hpp>
#include <flatfat_gpu.hpp>

struct res_t
{
    uint64_t key; uint64_t id; int64_t isum; double fsum;
    __host__ __device__ res_t(): key(0), id(0), isum(0), fsum(0.0) {}
    __host__ __device__ res_t(uint64_t _key, uint64_t _id): key(_key), id(_id), isum(0), fsum(0.0) {}
};

struct Comb
{
    __host__ __device__ void operator()(const res_t &a, const res_t &b, res_t &out) const
    {
        int64_t is = a.isum + b.isum; double fs = a.fsum + b.fsum;
        out.isum = is; out.fsum = fs;
    }
};

using fat_t = wf::FlatFAT_GPU<res_t, uint64_t, Comb>;
using item_t = wf::batch_item_gpu_t<res_t>;

struct RefKey
{
    fat_t *fat;
    uint64_t next_gwid, count, count_triggerer;
};

struct RefFfatGpu
{
    uint64_t W, S, Nb, B;
    int numSMs, maxBlocks;
    cudaStream_t stream;
    item_t *d_out; // Nb items
    RefKey key;    // single key (the per-key loop of :782-800 is host code; one key is what a tree sees)
};

extern "C" {

/* One key of a count-based Ffat_Windows_GPU (Key_Descriptor, ffat_replica_gpu.hpp:438-506). */
void *wfref_ffat_gpu_create(uint64_t win, uint64_t slide, uint64_t nb, uint64_t key)
{
    RefFfatGpu *h = new RefFfatGpu();
    h->W = win; h->S = slide; h->Nb = nb; h->B = (nb - 1) * slide + win;
    int dev; cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&h->numSMs, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&h->maxBlocks, cudaDevAttrMaxBlocksPerMultiprocessor, dev);
    cudaStreamCreate(&h->stream);
    cudaMalloc(&h->d_out, sizeof(item_t) * nb);
    h->key.fat = new fat_t(Comb(), h->B, nb, win, slide, key, h->numSMs, h->maxBlocks);
    h->key.next_gwid = 0; h->key.count = 0; h->key.count_triggerer = h->B;
    return h;
}

void wfref_ffat_gpu_destroy(void *hh)
{
    RefFfatGpu *h = reinterpret_cast<RefFfatGpu *>(hh);
    delete h->key.fat;
    cudaFree(h->d_out);
    cudaStreamDestroy(h->stream);
    delete h;
}

/* process_wins_cb for `num` lifted results of this key living in DEVICE memory (d_res). Window results are
 * copied to host arrays out/out_ts (Nb per trigger). Returns the number of results. */
uint64_t wfref_ffat_gpu_process(void *hh, res_t *d_res, uint64_t num, uint64_t wm,
                                res_t *out, uint64_t *out_ts, uint64_t out_cap)
{
    RefFfatGpu *h = reinterpret_cast<RefFfatGpu *>(hh);
    RefKey &k = h->key;
    uint64_t nout = 0, off = 0;
    item_t *tmp = (item_t *) malloc(sizeof(item_t) * h->Nb);
    while (k.count + num >= k.count_triggerer) {
        uint64_t take = k.count_triggerer - k.count;
        k.fat->add_cb(d_res + off, take, h->stream);
        num -= take; off += take; k.count += take;
        if (k.count_triggerer == h->B) k.fat->build(h->stream);
        else k.fat->update(h->S * h->Nb, h->stream);
        k.fat->computeResults(h->d_out, k.next_gwid, wm, h->stream);
        cudaMemcpy(tmp, h->d_out, sizeof(item_t) * h->Nb, cudaMemcpyDeviceToHost);
        for (uint64_t i = 0; i < h->Nb; i++) {
            if (nout < out_cap) { out[nout] = tmp[i].tuple; out_ts[nout] = tmp[i].timestamp; }
            nout++;
        }
        k.next_gwid += h->Nb;
        k.count_triggerer += h->S * h->Nb;
    }
    if (num > 0) { k.fat->add_cb(d_res + off, num, h->stream); k.count += num; }
    cudaStreamSynchronize(h->stream);
    free(tmp);
    return nout;
}

} // extern "C"
