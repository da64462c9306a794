"""world_size-2 runs of the CUDA path (SURVEY 8e, Appendix D "N-GPU vs 1-GPU meta-step"): two processes, each driving
the kernels through the C ABI.  With two or more GPUs visible the ranks sit on cuda:0 / cuda:1 and talk NCCL; on a
one-GPU box both ranks share cuda:0 and the single collective of the path goes through gloo's CUDA-tensor all-reduce, so
that the world > 1 branches of FOMetaMixin / MultiMixin are executed on hardware either way."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from tests.helpers import GOLD, check_adam_weights, load_batch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        return sk.getsockname()[1]


def _init(rank, world, port_no):
    two = torch.cuda.device_count() >= 2
    os.environ.update({"RANK": str(rank), "WORLD_SIZE": str(world), "LOCAL_RANK": str(rank if two else 0),
                       "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port_no)})
    from metaasr_crossaccent_b200 import dist as D
    D.init_from_env("nccl" if two else "gloo")
    return D


def _fomaml_worker(rank, world, port_no, out):
    from tests.test_e2e_gpu import load_tiny, make_solver
    D = _init(rank, world, port_no)
    z = np.load(GOLD / "fomaml_tiny.npz")
    s = make_solver("fomaml")
    load_tiny(s)
    tasks = []
    for acc in D.partition_tasks([0, 1], 2):
        tr = [(acc, load_batch(z, f"s0.a{acc}.tr{j}.")) for j in range(2)]
        tasks.append((tr, (acc, load_batch(z, f"s0.a{acc}.te."))))
    s.meta_step_on_tasks(tasks, global_task_count=2)
    osd = s.optimizer_state()                  # (a collective when the Adam moments are sharded: NVLS meta-update)
    torch.save({"w": s._original_flat.cpu(), "cs": s.replica_checksum().cpu(), "m": osd["m"], "v": osd["v"],
                "nvls": getattr(s, "_nvls", None) is not None, "upd_absmax": float(s._upd_flat.abs().max())},
               f"{out}/w{rank}.pt")
    torch.distributed.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_rank_fomaml_meta_step_on_cuda_equals_sequential_and_golden(tmp_path):
    """Accents partitioned over 2 ranks + ONE all-reduce of the flat update arena on the CUDA path: replicas end
    bit-identical, equal the 1-rank sequential meta-step (rel-L2 <= 1e-6 on the meta weights' change is not measurable
    through Adam, so the comparison is in units of lr) and the live-reference golden of the same meta-step."""
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from tests.test_e2e_gpu import load_tiny, make_solver
    mp.spawn(_fomaml_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    r0, r1 = torch.load(tmp_path / "w0.pt"), torch.load(tmp_path / "w1.pt")
    assert torch.equal(r0["w"], r1["w"]) and torch.equal(r0["cs"], r1["cs"])          # replicas in sync, bit for bit
    z = np.load(GOLD / "fomaml_tiny.npz")
    s = make_solver("fomaml")
    load_tiny(s)
    tasks = []
    for acc in range(2):
        tr = [(acc, load_batch(z, f"s0.a{acc}.tr{j}.")) for j in range(2)]
        tasks.append((tr, (acc, load_batch(z, f"s0.a{acc}.te."))))
    s.meta_step_on_tasks(tasks)
    lr = s.meta_opt.lr
    eng = s.asr_model.engine
    w0 = r0["w"]
    for n in eng.layout.offsets:
        check_adam_weights(z, "s0.w.", ["s0.mg."], n, eng.layout.view(w0, n), lr)
    d = (w0 - s._original_flat.cpu()).abs()
    assert float(d.max()) <= 2.0 * lr + 1e-12
    assert float((d > 3e-2 * lr).float().mean()) < 0.02
    # two GPUs under NCCL on an NVSwitch box: the outer update ran as ONE kernel over the multicast mappings
    # (masr_nvls_reduce_adam) with the Adam moments sharded over the ranks; gathered, they equal the 1-rank moments, and
    # the kernel left every rank's update arena cleared
    if torch.cuda.device_count() >= 2:
        print("NVLS meta-update:", r0["nvls"], r1["nvls"])
    assert r0["nvls"] == r1["nvls"]
    assert r0["upd_absmax"] == 0.0 and r1["upd_absmax"] == 0.0
    st = s.meta_opt.state
    for key, ref in (("m", st.m.cpu()), ("v", st.v.cpu())):
        assert torch.equal(r0[key], r1[key])
        assert float((r0[key] - ref).norm()) <= 1e-4 * float(ref.norm()) + 1e-12, key


def _multi_worker(rank, world, port_no, out):
    from tests.test_e2e_gpu import load_tiny, make_solver
    _init(rank, world, port_no)
    z = np.load(GOLD / "multi_tiny.npz")
    s = make_solver("multi", meta=False)
    load_tiny(s)
    losses = []
    for step in range(2):                      # rank r trains on batch s{2*step + r} of the golden file's inputs
        losses.append(s.multi_step((0, load_batch(z, f"s{(2 * step + rank) % 3}.")))["loss"])
    torch.save({"w": s.asr_model.engine.params.cpu(), "loss": losses, "lr": s.asr_opt.lr,
                "step_num": s.asr_opt.step_num}, f"{out}/m{rank}.pt")
    torch.distributed.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_rank_multi_task_gradient_dp_on_cuda(tmp_path):
    """Multi-task data parallelism (multi_interface.py:100-114) on the CUDA path: ranks draw different
    batches, all-reduce the gradient arena, clip + noam-Adam on the MEAN gradient.  Replicas end bit-identical and equal
    one process stepping on the averaged gradient of the same two batches."""
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from metaasr_crossaccent_b200 import interfaces as I
    from tests.test_e2e_gpu import load_tiny, make_solver
    mp.spawn(_multi_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    r0, r1 = torch.load(tmp_path / "m0.pt"), torch.load(tmp_path / "m1.pt")
    assert torch.equal(r0["w"], r1["w"]) and r0["lr"] == r1["lr"] and r0["step_num"] == r1["step_num"] == 2
    z = np.load(GOLD / "multi_tiny.npz")
    s = make_solver("multi", meta=False)
    load_tiny(s)
    eng = s.asr_model.engine
    n = eng.layout.total
    for step in range(2):
        gsum = torch.zeros_like(eng.grads)
        for r, ref in ((0, r0), (1, r1)):
            info = s.run_batch(0, *load_batch(z, f"s{(2 * step + r) % 3}."), train=True, accent_idx=0)
            assert abs(info["loss"] - ref["loss"][step]) <= 1e-5 * abs(info["loss"])      # each rank saw its own batch
            gsum += eng.grads
        eng.grads.copy_(gsum / 2)
        eng.be.mt_sumsq(eng.grads[:n], s._gnorm)
        s.asr_opt.step(s._gnorm, I.GRAD_CLIP)
    assert abs(s.asr_opt.lr - r0["lr"]) < 1e-15
    diff = (eng.params.cpu() - r0["w"]).abs()                    # Adam(eps 1e-9): compare in units of lr
    assert float(diff.max()) <= 2.0 * s.asr_opt.lr + 1e-12
    assert float((diff > 3e-2 * s.asr_opt.lr).float().mean()) < 0.02
