"""Host-side (CPU) profile of issuing training steps: where the Python time of a step goes.

    python tools/host_profile.py [c2] [steps]      -> gpurun_out/host_profile_<tag>.txt

The step is issued without any synchronisation inside the profiled region, so the numbers are pure launch-path cost
(autograd, ctypes marshalling, torch op dispatch), to be compared with the device time of the same steps."""
import cProfile
import io
import pstats
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "c2"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    from mlx_cuda_distributed_pretraining_b200.core.training import Config, Trainer
    d = bench.load_config(tag, distributed=False)
    c = bench.dims_of(d)
    k = c["accum"]
    d["training"]["hyperparameters"]["iters"] = 1000
    tr = Trainer(Config.from_dict(d), synthetic=True, quiet=True, run_root=str(ROOT / "gpurun_out" / "bench_runs"))
    tr._accum_step, tr._accum_tokens = 0, 0
    n = (steps * 2 + 3) * k
    batches = [tr.data_manager.generate_batch(s).to(tr.device) for s in range(n)]
    it = iter(range(n))

    def run(nsteps):
        for _ in range(nsteps * k):
            i = next(it)
            tr.micro_step(i, batches[i])

    run(3)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t0 = time.perf_counter()
    run(steps)
    host = time.perf_counter() - t0
    e1.record()
    torch.cuda.synchronize()
    dev = e0.elapsed_time(e1)
    pr = cProfile.Profile()
    pr.enable()
    run(steps)
    pr.disable()
    torch.cuda.synchronize()
    out = io.StringIO()
    out.write(f"{tag}: {steps} steps, host issue {1e3 * host / steps:.2f} ms/step, device {dev / steps:.2f} ms/step\n")
    for key in ("tottime", "cumulative"):
        out.write(f"\n==== sorted by {key} ====\n")
        pstats.Stats(pr, stream=out).strip_dirs().sort_stats(key).print_stats(45)
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / f"host_profile_{tag}.txt").write_text(out.getvalue())
    print(out.getvalue().splitlines()[0])


if __name__ == "__main__":
    main()
