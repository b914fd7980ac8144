"""Key-sharded execution of the pipeline across the GPUs of one box (one process per GPU, torch.distributed).

The reference is single-GPU (README.md:15); its keyed operators shard by `std::hash<key_t>()(key) % num_dests`
(wf/keyby_emitter.hpp:215-217, wf/keyby_emitter_gpu.hpp:621) -- the identity hash for integer keys -- and that is the
rule used here across ranks. Map_GPU / Filter_GPU are stateless replicas; the window state partitions by key:

  global step t covers world*K consecutive batches of the stream; rank r owns the K batches [r*K, (r+1)*K) of that span,
  runs the fused Map->Filter over them, partitions the survivors into `world` shard segments (stable), exchanges the
  segment sizes and then the segments (all-to-all over NVLink), and feeds what it receives -- one chunk per source rank,
  in source-rank order, which is global stream order -- to its Ffat_Windows_GPU replica (keys with key % world == rank).

Because every rank receives its keys' items in global arrival order, every count window is the one a single operator
would produce. Only host-side plumbing lives here (testable on CPU with the gloo backend); kernels stay behind the C ABI.
"""
import torch
import torch.distributed as dist


def owner_span(step, rank, world, seg_tuples):
    """[first, last) stream indices of the segment rank `rank` ingests at global step `step`."""
    first = (step * world + rank) * seg_tuples
    return first, first + seg_tuples


def exchange_counts(send_counts, watermark=0):
    """send_counts[d] = items this rank sends to rank d  ->  (recv_counts[s], watermark[s]) = what rank s sends to this
    rank and the watermark of rank s's segment (both travel in one small all-to-all)."""
    meta = torch.stack([send_counts, torch.full_like(send_counts, int(watermark))], dim=1).contiguous()
    recv = torch.empty_like(meta)
    dist.all_to_all_single(recv, meta)
    return recv[:, 0].contiguous(), recv[:, 1].contiguous()


def exchange_segments(send_buf, send_counts_h, recv_counts_h, item_bytes, recv_buf=None):
    """All-to-all of variable-size shard segments. send_buf holds the segments for rank 0..world-1 back to back
    (uint8, item_bytes per item). Returns (recv_buf, offsets) where chunk s = recv_buf[offsets[s]*item_bytes :
    offsets[s+1]*item_bytes] came from source rank s, in its arrival order."""
    n_recv = int(sum(recv_counts_h))
    if recv_buf is None or recv_buf.numel() < n_recv * item_bytes:
        recv_buf = torch.empty(max(1, n_recv) * item_bytes, dtype=torch.uint8, device=send_buf.device)
    n_send = int(sum(send_counts_h))
    dist.all_to_all_single(recv_buf[:n_recv * item_bytes], send_buf[:n_send * item_bytes],
                           output_split_sizes=[int(c) * item_bytes for c in recv_counts_h],
                           input_split_sizes=[int(c) * item_bytes for c in send_counts_h])
    offs = [0]
    for c in recv_counts_h:
        offs.append(offs[-1] + int(c))
    return recv_buf, offs


def exchange_regions(regions, region_cap, send_counts_h, recv_counts_h, item_bytes, recv_buf=None):
    """All-to-all of the shard regions written by wfb_shard_lift (region d starts at d*region_cap items). Returns
    (recv_buf, offsets): chunk s came from source rank s, in its arrival order."""
    world = len(send_counts_h)
    n_recv = int(sum(recv_counts_h))
    if recv_buf is None or recv_buf.numel() < max(1, n_recv) * item_bytes:
        recv_buf = torch.empty(max(1, n_recv) * item_bytes, dtype=torch.uint8, device=regions.device)
    offs = [0]
    for c in recv_counts_h:
        offs.append(offs[-1] + int(c))
    ins = [regions[d * region_cap * item_bytes:(d * region_cap + int(send_counts_h[d])) * item_bytes] for d in range(world)]
    outs = [recv_buf[offs[s] * item_bytes:offs[s + 1] * item_bytes] for s in range(world)]
    dist.all_to_all(outs, ins)
    return recv_buf, offs


class KeyShardedPipeline:
    """Map_GPU -> Filter_GPU -> (keyby across GPUs) -> Ffat_Windows_GPU on this rank's key shard.

    Source side: ONE fused pass (wfb_shard_lift): map, filter, lift and the stable partition of the 32-byte lifted
    results by key % world. Exchange: sizes + watermarks in one small all-to-all, then the records (NCCL, NVLink).
    Destination side: the window operator instantiated for already-lifted records (WFB_PROG_LIFTED32)."""

    def __init__(self, ops, functors, win, slide, nb, max_keys, rank, world, device, pipelined=True):
        self.ops, self.f, self.rank, self.world, self.dev = ops, functors, rank, world, device
        self.eng = ops.Engine(ops.PROG_TUPLE64)
        self.ff = ops.FfatWindowsGPU(ops.PROG_LIFTED32, win, slide, nb, max_keys=max_keys, dense_keys=True, pipelined=pipelined)
        self.rb = self.eng.result_bytes
        self.regions = self.recv = None
        self.region_cap = 0
        self.counts = torch.zeros(9, dtype=torch.int32, device=device)

    def _ensure(self, n):
        if self.region_cap < n:
            self.region_cap = n  # worst case: every item of the segment survives and goes to one shard
            self.regions = torch.empty(self.world * n * self.rb, dtype=torch.uint8, device=self.dev)

    def step(self, batches, watermark, out, out_ts, n_out):
        """batches: this rank's K batches of the global step. Results of the window operator go to out."""
        ops = self.ops
        self._ensure(sum(b.n for b in batches))
        self.eng.shard_lift(batches, self.f, self.world, self.regions, self.region_cap, self.counts)
        cnt_h = self.counts.cpu().tolist()                         # the only host sync of the source side
        if cnt_h[8]:
            raise RuntimeError("wfb_shard_lift: shard region overflow")
        send_counts_h = cnt_h[:self.world]
        send_counts = torch.tensor(send_counts_h, dtype=torch.int64, device=self.dev)
        rc, rw = exchange_counts(send_counts, watermark)
        recv_counts_h, wms = rc.cpu().tolist(), rw.cpu().tolist()
        self.recv, offs = exchange_regions(self.regions, self.region_cap, send_counts_h, recv_counts_h, self.rb, self.recv)
        chunks = [ops.DeviceBatch(self.recv[offs[s] * self.rb:offs[s + 1] * self.rb], None, offs[s + 1] - offs[s], wms[s])
                  for s in range(self.world)]
        self.ff.process(chunks, pre=None, out=out, out_ts=out_ts, n_out=n_out)
        return sum(recv_counts_h)
