"""What the PCIe link of this box sustains: host->device alone, device->host alone, both at once (two streams,
pinned memory), for a few transfer sizes.  Context for bench.py's e2e number, which moves ~100 MB in and ~113 MB out
per step.  Usage (on the GPU box): python tools/gpu_pcie.py"""
import time
import torch


def run(nbytes_in, nbytes_out, chunks=1, reps=10):
    hi = torch.empty(nbytes_in, dtype=torch.uint8).pin_memory()
    ho = torch.empty(nbytes_out, dtype=torch.uint8).pin_memory()
    di = torch.empty(nbytes_in, dtype=torch.uint8, device="cuda")
    do = torch.empty(nbytes_out, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def go(do_in, do_out):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            ci, co = nbytes_in // chunks, nbytes_out // chunks
            for k in range(chunks):
                if do_in:
                    with torch.cuda.stream(s1):
                        di[k * ci:(k + 1) * ci].copy_(hi[k * ci:(k + 1) * ci], non_blocking=True)
                if do_out:
                    with torch.cuda.stream(s2):
                        ho[k * co:(k + 1) * co].copy_(do[k * co:(k + 1) * co], non_blocking=True)
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps

    go(True, True)
    a, b, c = go(True, False), go(False, True), go(True, True)
    print(f"in {nbytes_in/1e6:.0f} MB out {nbytes_out/1e6:.0f} MB chunks {chunks}: h2d alone {nbytes_in/a/1e9:.1f} GB/s, "
          f"d2h alone {nbytes_out/b/1e9:.1f} GB/s, both {c*1e3:.3f} ms = {(nbytes_in+nbytes_out)/c/1e9:.1f} GB/s total "
          f"(in {nbytes_in/c/1e9:.1f} + out {nbytes_out/c/1e9:.1f})")


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    run(100_000_000, 112_000_000, 1)
    run(100_000_000, 112_000_000, 8)
    run(400_000_000, 448_000_000, 1)
    run(100_000_000, 100_000_000, 1)


def run_with_kernel_load(nbytes_in, nbytes_out, chunks=8, reps=6):
    """The same duplex transfer while a memory-bound kernel (x += 1 over 1 GiB) keeps the SMs and HBM busy."""
    hi = torch.empty(nbytes_in, dtype=torch.uint8).pin_memory()
    ho = torch.empty(nbytes_out, dtype=torch.uint8).pin_memory()
    di = torch.empty(nbytes_in, dtype=torch.uint8, device="cuda")
    do = torch.empty(nbytes_out, dtype=torch.uint8, device="cuda")
    big = torch.zeros(1 << 28, dtype=torch.float32, device="cuda")
    s1, s2, s3 = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
    for load in (False, True):
        torch.cuda.synchronize()
        if load:
            with torch.cuda.stream(s3):
                for _ in range(400):
                    big.add_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ci, co = nbytes_in // chunks, nbytes_out // chunks
        e0.record(s1)
        s2.wait_event(e0)
        for _ in range(reps):
            for k in range(chunks):
                with torch.cuda.stream(s1):
                    di[k * ci:(k + 1) * ci].copy_(hi[k * ci:(k + 1) * ci], non_blocking=True)
                with torch.cuda.stream(s2):
                    ho[k * co:(k + 1) * co].copy_(do[k * co:(k + 1) * co], non_blocking=True)
        ev2 = torch.cuda.Event(); ev2.record(s2)
        s1.wait_event(ev2)
        e1.record(s1)
        e1.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(f"duplex {chunks} chunks, kernel load {load}: {ms:.3f} ms per {nbytes_in/1e6:.0f}+{nbytes_out/1e6:.0f} MB = "
              f"{(nbytes_in+nbytes_out)/ms/1e6:.1f} GB/s total")
        torch.cuda.synchronize()


if __name__ == "__main__":
    run_with_kernel_load(100_000_000, 112_000_000, 8)
    run_with_kernel_load(100_000_000, 112_000_000, 1)
