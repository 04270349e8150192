"""Worker of tests/test_gpu_parity_wide.py::test_two_nccl_ranks_equal_one_gpu (launched with torch.distributed.run,
one rank per GPU): reconstruct_sharded over NCCL must equal the single-GPU call bit for bit."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from oracle import defensegan_oracle as O
    from defensegan_b200.models.gan import MnistDefenseGAN
    from defensegan_b200.parallel import reconstruct_sharded
    res = {}
    for precision in ("fp16", "fp32"):
        gan = MnistDefenseGAN(test_mode=True, verbose=False, precision=precision)
        gan.rec_rr, gan.rec_iters = 4, 8
        for name, B in (("", 16), ("_ragged", 13)):
            x = torch.tensor(O.synthetic_images("mnist", gan.weights, B)).to(dev)
            z0 = torch.tensor(O.sample_z0(B * 4, 128)).to(dev)
            single = gan.reconstruct(x, z_init_val=z0)
            sharded = reconstruct_sharded(gan, x, z_init_val=z0)
            key = "equal_ragged" if name else "equal_given_z0"
            res[key] = res.get(key, True) and bool(torch.equal(single, sharded))
            # shared Philox stream: same counter state before both calls
            c0 = gan._call_counter
            single_r = gan.reconstruct(x)
            gan._call_counter = c0
            sharded_r = reconstruct_sharded(gan, x)
            res["equal_random_z0"] = res.get("equal_random_z0", True) and bool(torch.equal(single_r, sharded_r))
        gan.close()
    flags = torch.tensor([int(v) for v in res.values()], device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    if rank == 0:
        with open(sys.argv[1], "w") as f:
            json.dump({k: bool(v) for k, v in zip(res.keys(), flags.tolist())}, f)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
