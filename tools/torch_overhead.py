"""Host cost of tiny torch operations on this box (they sit on the per-step path of windflow_b200/multigpu.py)."""
import time, torch
torch.cuda.set_device(0)
a = torch.zeros(8, dtype=torch.int32, device="cuda"); b = torch.ones(8, dtype=torch.int32, device="cuda")
c = torch.zeros(8, 2, dtype=torch.int64, device="cuda")
h = torch.zeros(8, dtype=torch.int32).pin_memory()
ev = torch.cuda.Event()
s2 = torch.cuda.Stream()
def t(name, f, n=2000):
    torch.cuda.synchronize(); f()
    t0 = time.perf_counter()
    for _ in range(n): f()
    dt = (time.perf_counter() - t0) / n * 1e6
    torch.cuda.synchronize()
    print(f"{name:40s} {dt:8.1f} us/op")
t("a.add_(b)", lambda: a.add_(b))
t("c[:,0].copy_(a)", lambda: c[:, 0].copy_(a))
t("c[:,1].fill_(7)", lambda: c[:, 1].fill_(7))
t("h.copy_(a, non_blocking=True)", lambda: h.copy_(a, non_blocking=True))
t("ev.record()", lambda: ev.record())
t("s2.wait_event(ev)", lambda: s2.wait_event(ev))
def ctx():
    with torch.cuda.stream(s2):
        pass
t("with torch.cuda.stream(s2)", ctx)
t("torch.cuda.current_stream()", lambda: torch.cuda.current_stream())
