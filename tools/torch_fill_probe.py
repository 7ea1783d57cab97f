"""torch's fill_ / copy_ on the same sizes as tools/write_probe.cu (the round-1 'write-only ceiling' came from fill_)."""
import json
import torch

n = (486 << 20)
buf = torch.empty(n, dtype=torch.int32, device="cuda")
src = torch.empty(n // 2, dtype=torch.int32, device="cuda")
out = {}
def best(f, reps=7):
    b = 1e9
    for i in range(reps):
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(); f(); a1.record(); a1.synchronize()
        if i:
            b = min(b, a0.elapsed_time(a1))
    return b
ms = best(lambda: buf.fill_(7))
out["torch_fill"] = {"ms": ms, "gbs": 4 * n / ms / 1e6}
ms = best(lambda: buf.zero_())
out["torch_zero"] = {"ms": ms, "gbs": 4 * n / ms / 1e6}
ms = best(lambda: buf[: n // 2].copy_(src))
out["torch_copy_half"] = {"ms": ms, "gbs_rw": 4 * n / ms / 1e6}
print(json.dumps(out))
