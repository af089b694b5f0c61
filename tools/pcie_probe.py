"""Host <-> device copy bandwidth, one GPU alone and all visible GPUs at once (profiling aid: is the host link shared?)."""
import threading, time, torch


def run(dev_ids, mb=20, reps=20):
    res = {}
    bar = threading.Barrier(len(dev_ids))

    def work(i):
        dev = torch.device("cuda", i)
        n = mb << 20
        h = torch.empty(n, dtype=torch.uint8, pin_memory=True); h.fill_(3)
        d = torch.empty(n, dtype=torch.uint8, device=dev)
        s = torch.cuda.Stream(device=dev)
        with torch.cuda.stream(s):
            for _ in range(3): d.copy_(h, non_blocking=True)
            s.synchronize()
            bar.wait()
            t0 = time.perf_counter()
            for _ in range(reps): d.copy_(h, non_blocking=True)
            s.synchronize()
            res[i] = n * reps / (time.perf_counter() - t0) / 1e9
    th = [threading.Thread(target=work, args=(i,)) for i in dev_ids]
    for t in th: t.start()
    for t in th: t.join()
    return res


n = torch.cuda.device_count()
print("alone:", {k: round(v, 1) for k, v in run([0]).items()}, "GB/s H2D (20 MiB copies)")
if n > 1:
    r = run(list(range(n)))
    print(f"{n} GPUs at once:", {k: round(v, 1) for k, v in r.items()}, "sum", round(sum(r.values()), 1))
