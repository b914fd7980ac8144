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


class KeyShardedPipeline:
    """Map_GPU -> Filter_GPU -> (keyby across GPUs) -> Ffat_Windows_GPU on this rank's key shard."""

    def __init__(self, ops, prog, functors, win, slide, nb, max_keys, rank, world, device, pipelined=True):
        self.ops, self.f, self.rank, self.world, self.dev = ops, functors, rank, world, device
        self.eng = ops.Engine(prog)
        self.ff = ops.FfatWindowsGPU(prog, win, slide, nb, max_keys=max_keys, dense_keys=True, pipelined=pipelined)
        self.tb = self.eng.tuple_bytes
        self.filt = self.part = self.recv = None
        self.n_f = torch.zeros(1, dtype=torch.int32, device=device)
        self.launches_extra = 0

    def _ensure(self, n):
        if self.filt is None or self.filt.numel() < n * self.tb:
            self.filt = torch.empty(n * self.tb, dtype=torch.uint8, device=self.dev)
            self.part = torch.empty(n * self.tb, dtype=torch.uint8, device=self.dev)

    def step(self, seg, out, out_ts, n_out):
        """seg: one DeviceBatch holding this rank's K batches back to back. Results of the window operator go to out."""
        ops = self.ops
        self._ensure(seg.n)
        fb = ops.DeviceBatch(self.filt, None, seg.n, seg.watermark)
        inb = ops.DeviceBatch(seg.tuples, None, seg.n, seg.watermark)
        self.eng.map_filter(inb, self.f, out=fb, n_out=self.n_f)                   # fused Map -> Filter over the segment
        n = int(self.n_f.item())
        fb.n = n
        pb = ops.DeviceBatch(self.part, None, n, seg.watermark)
        _, seg_off = self.eng.shard_by_key(fb, self.world, out=pb)                  # stable partition by key % world
        off_h = seg_off.cpu().tolist()
        send_counts_h = [off_h[d + 1] - off_h[d] for d in range(self.world)]
        send_counts = torch.tensor(send_counts_h, dtype=torch.int64, device=self.dev)
        rc, rw = exchange_counts(send_counts, seg.watermark)
        recv_counts_h, wms = rc.cpu().tolist(), rw.cpu().tolist()
        self.recv, offs = exchange_segments(self.part, send_counts_h, recv_counts_h, self.tb, self.recv)
        batches = [ops.DeviceBatch(self.recv[offs[s] * self.tb:offs[s + 1] * self.tb], None, offs[s + 1] - offs[s], wms[s])
                   for s in range(self.world)]
        self.ff.process(batches, pre=None, out=out, out_ts=out_ts, n_out=n_out)
        return sum(recv_counts_h)
