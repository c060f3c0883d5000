"""Generates tests/golden/randperm_cuda.pt ON A GPU BOX: permutations drawn by torch's own
`torch.randperm(n, device="cuda")` together with the generator state (seed, offset) they were
drawn from. The CPU oracle (oracle/randperm_ref.py) and the CUDA mask kernel are pinned to these.

    gpurun -- python oracle/make_randperm_golden.py gpurun_out/randperm_cuda.pt
"""
import sys

import torch


def main(out_path):
    dev = torch.cuda.current_device()
    gen = torch.cuda.default_generators[dev]
    cases = []
    for seed, n, reps, pre in [(42, 196, 64, 0), (7, 196, 300, 12), (1234, 36, 40, 4), (5, 144, 20, 0),
                               (99, 1, 3, 0), (100, 2, 8, 0), (3, 257, 4, 8), (11, 576, 3, 0), (2 ** 40 + 17, 1024, 2, 0)]:
        torch.manual_seed(seed)
        if pre:
            torch.rand(pre, device="cuda")
        seed_, off0 = gen.initial_seed(), gen.get_offset()
        perms = torch.stack([torch.randperm(n, device="cuda") for _ in range(reps)]).cpu()
        cases.append(dict(seed=int(seed_), offset=int(off0), n=n, reps=reps, perms=perms.to(torch.int16 if n < 32768 else torch.int64),
                          offset_after=int(gen.get_offset())))
    torch.save(dict(torch=str(torch.__version__), device=torch.cuda.get_device_name(dev), cases=cases), out_path)
    print("wrote", out_path, "cases", len(cases))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "tests/golden/randperm_cuda.pt")
