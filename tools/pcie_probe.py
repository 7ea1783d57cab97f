"""PCIe floor of the e2e tick: pinned H2D, D2H and both at once, at the tick's transfer sizes (torch copies, CUDA events).
Run on the GPU box:  python tools/pcie_probe.py   (prints one JSON line; numbers quoted in DESIGN.md / profiles/)"""
import json
import time

import torch


def main():
    dev = torch.device("cuda:0")
    h2d_bytes, d2h_bytes = 18_690_704, 8_516_688
    hs = torch.empty(h2d_bytes, dtype=torch.uint8).pin_memory()
    hd = torch.empty(d2h_bytes, dtype=torch.uint8).pin_memory()
    ds = torch.empty(h2d_bytes, dtype=torch.uint8, device=dev)
    dd = torch.empty(d2h_bytes, dtype=torch.uint8, device=dev)
    s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()

    def run(up, dn, reps=20):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            if up:
                with torch.cuda.stream(s_up):
                    ds.copy_(hs, non_blocking=True)
            if dn:
                with torch.cuda.stream(s_dn):
                    hd.copy_(dd, non_blocking=True)
            s_up.synchronize()
            s_dn.synchronize()
        return (time.perf_counter() - t0) / reps * 1e3

    for _ in range(2):
        run(True, True, 3)
    out = {"h2d_ms": run(True, False), "d2h_ms": run(False, True), "both_ms": run(True, True)}
    out["h2d_gbs"] = h2d_bytes / out["h2d_ms"] / 1e6
    out["d2h_gbs"] = d2h_bytes / out["d2h_ms"] / 1e6
    out["both_gbs"] = (h2d_bytes + d2h_bytes) / out["both_ms"] / 1e6
    # small-chunk D2H like chd_fetch_results issues (12 copies)
    chunks = [400_004, 437_888, 437_888, 437_888, 110_000, 110_000, 110_000, 110_000, 5_186_832, 110_892, 110_892, 110_892, 400_000, 800_008]
    offs, o = [], 0
    for c in chunks:
        offs.append(o)
        o += c
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        with torch.cuda.stream(s_dn):
            for c, of in zip(chunks, offs):
                hd[of:of + c].copy_(dd[of:of + c], non_blocking=True)
        s_dn.synchronize()
    out["d2h_chunked_ms"] = (time.perf_counter() - t0) / 20 * 1e3
    print(json.dumps(out))


if __name__ == "__main__":
    main()
