"""Worker for the multi-process tests (launched by tests/test_dist.py).

mode gloo : CPU, world 2.  Host-side DP logic: shard bounds, unique-id style byte broadcast, and the identity the
            GPU path relies on — mean over ranks of per-shard gradients (local BN statistics) == oracle computed
            shard-wise and averaged; a data-parallel SGD-momentum step on the averaged gradient keeps replicas identical.
mode nccl : GPU, one rank per GPU.  The CUDA path with its NCCL all-reduce vs the same oracle average.
mode syncbn : GPU, one rank per GPU.  tcr_comm_set_sync_bn: every rank's step must equal the oracle's SINGLE-device step on the
            global batch (gradient, updated parameters, momentum slots, BatchNorm moving statistics).
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import tcresnet_b200  # noqa: E402,F401
from tcresnet_b200.dp import broadcast_bytes, shard_bounds  # noqa: E402
from oracle import tcr_oracle as O  # noqa: E402
from parity_cases import perturbed_variables  # noqa: E402


def oracle_shard_grads(spec, params, moving, feat, onehot, wd):
    logits, cache = O.forward(spec, params, moving, feat, True)
    return O.flatten_vars(spec, O.backward(spec, params, cache, logits, onehot, wd), np.float64)


def main():
    mode = sys.argv[1]
    if mode == "gloo":
        dist.init_process_group("gloo")
    else:
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)                      # before ANY NCCL collective (one GPU per rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()

    def mark(msg):
        print(f"[rank {rank}] {msg}", file=sys.stderr, flush=True)
    n_global = 16 * world
    spec = O.build_spec("TCResNet8", 1.0, 49)
    params, moving = perturbed_variables(spec)
    wav, onehot = O.synthetic_batch(n_global)
    lo, hi = shard_bounds(n_global, rank, world)
    assert (hi - lo) * world == n_global and lo == rank * 16
    token = broadcast_bytes(bytes(range(128)) if rank == 0 else None)
    assert token == bytes(range(128))
    feat = O.mfcc(wav, 640, 320)
    wd, lr, mom = 1e-3, 0.1, 0.9
    ref = np.mean([oracle_shard_grads(spec, params, moving, feat[r * 16:(r + 1) * 16], onehot[r * 16:(r + 1) * 16], wd)
                   for r in range(world)], axis=0)
    if mode == "gloo":
        mine = torch.from_numpy(oracle_shard_grads(spec, params, moving, feat[lo:hi], onehot[lo:hi], wd))
        dist.all_reduce(mine)
        avg = (mine / world).numpy()
        assert np.abs(avg - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max())
        new = O.flatten_vars(spec, params, np.float64) - lr * avg          # slots start at 0
        gathered = [torch.zeros_like(torch.from_numpy(new)) for _ in range(world)]
        dist.all_gather(gathered, torch.from_numpy(new))
        assert all(torch.equal(g, gathered[0]) for g in gathered)          # replicas stay identical
    else:
        from tcresnet_b200.engine import Engine
        eng = Engine(max_batch=16, dropout_keep_prob=1.0)
        mark("engine created")
        eng.attach_process_group()
        mark("communicator attached")
        dev = eng.device
        p = torch.from_numpy(O.flatten_vars(spec, params)).to(dev)
        mv = torch.from_numpy(O.flatten_moving(spec, moving)).to(dev)
        sl = torch.zeros_like(p)
        out = eng.train_step(torch.from_numpy(wav[lo:hi]).to(dev), torch.from_numpy(onehot[lo:hi]).to(dev), p, sl, mv,
                             lr, mom, wd, want_grads=True)
        torch.cuda.synchronize()
        mark("train step done")
        g = out["grads"].cpu().numpy().astype(np.float64)
        err = np.abs(g - ref).max() / np.abs(ref).max()
        assert err < 1e-4, f"rank {rank}: averaged gradient off by {err}"
        expect = O.flatten_vars(spec, params, np.float64) - lr * ref
        perr = np.abs(p.cpu().numpy() - expect).max() / np.abs(expect).max()
        assert perr < 1e-4, perr
        for step in range(1, 4):                      # more steps: the peer-memory exchange alternates two gradient buffers
            eng.train_step(torch.from_numpy(wav[lo:hi]).to(dev), torch.from_numpy(onehot[lo:hi]).to(dev), p, sl, mv, lr, mom, wd,
                           dropout_seed=step)
        torch.cuda.synchronize()
        gathered = [torch.zeros_like(p) for _ in range(world)]
        dist.all_gather(gathered, p)
        assert all(torch.equal(x, gathered[0]) for x in gathered), "replicas diverged"
        assert torch.isfinite(p).all()
        if rank == 0:
            print(f"exchange={eng.exchange} world={world}: grad rel err {err:.2e}, params rel err {perr:.2e}, replicas bit-identical "
                  f"after 4 steps")
    dist.barrier()
    dist.destroy_process_group()


def syncbn_main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    per = 12
    n_global = per * world
    spec = O.build_spec("TCResNet8", 1.0, 49)
    params, moving = perturbed_variables(spec)
    rng = np.random.RandomState(3)
    slots = {k: 0.01 * rng.randn(*v.shape) for k, v in params.items()}
    wav, onehot = O.synthetic_batch(n_global)
    feat = O.mfcc(wav, 640, 320)
    wd, lr, mom = 1e-3, 0.1, 0.9
    # the reference's single device with the whole batch: global BatchNorm statistics, mean loss over world * per utterances
    p1, mv1, sl1, ref = O.train_step(spec, params, moving, slots, feat, onehot, lr, mom, wd, 1.0, None, 0.0)
    from tcresnet_b200.engine import Engine
    eng = Engine(max_batch=per, dropout_keep_prob=1.0)
    eng.attach_process_group()
    eng.set_sync_bn(True)
    dev = eng.device
    p = torch.from_numpy(O.flatten_vars(spec, params)).to(dev)
    mv = torch.from_numpy(O.flatten_moving(spec, moving)).to(dev)
    sl = torch.from_numpy(O.flatten_vars(spec, slots)).to(dev)
    lo, hi = rank * per, (rank + 1) * per
    out = eng.train_step(torch.from_numpy(wav[lo:hi]).to(dev), torch.from_numpy(onehot[lo:hi]).to(dev), p, sl, mv, lr, mom, wd,
                         want_grads=True)
    torch.cuda.synchronize()

    def rel(a, b):
        b = np.asarray(b, np.float64)
        return float(np.abs(np.asarray(a, np.float64) - b).max() / np.abs(b).max())
    errs = {"grads": rel(out["grads"].cpu().numpy(), O.flatten_vars(spec, ref["grads"], np.float64)),
            "params": rel(p.cpu().numpy(), O.flatten_vars(spec, p1, np.float64)),
            "slots": rel(sl.cpu().numpy(), O.flatten_vars(spec, sl1, np.float64)),
            "moving": rel(mv.cpu().numpy(), O.flatten_moving(spec, mv1, np.float64))}
    assert all(v < 1e-4 for v in errs.values()), f"rank {rank}: {errs}"
    # and it differs from local statistics: the flag is not a no-op
    eng.set_sync_bn(False)
    p2 = torch.from_numpy(O.flatten_vars(spec, params)).to(dev)
    out2 = eng.train_step(torch.from_numpy(wav[lo:hi]).to(dev), torch.from_numpy(onehot[lo:hi]).to(dev), p2, torch.zeros_like(p2),
                          torch.from_numpy(O.flatten_moving(spec, moving)).to(dev), lr, mom, wd, want_grads=True)
    local_err = rel(out2["grads"].cpu().numpy(), O.flatten_vars(spec, ref["grads"], np.float64))
    assert local_err > 10 * errs["grads"], (local_err, errs)
    if rank == 0:
        print(f"syncbn world={world}: vs the oracle's single-device batch of {n_global}: " + ", ".join(f"{k} {v:.2e}" for k, v in errs.items()) +
              f"; local-statistics gradient differs by {local_err:.2e}")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    if sys.argv[1] == "syncbn":
        syncbn_main()
    else:
        main()
