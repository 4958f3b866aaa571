"""Checkpoint interop: neutral network / memory files next to the Julia session files."""
import json
import os

import numpy as np
import pytest

import _pkg


@pytest.fixture(scope="module")
def ck():
    _pkg.load()
    import alphazero_jl_b200.checkpoint as m
    return m


def _samples(k, seed=0):
    rng = np.random.default_rng(seed)
    return dict(states=rng.integers(0, 3, (k, 43)).astype(np.uint8), pi=rng.random((k, 7)), z=rng.normal(size=k), t=rng.integers(1, 43, k).astype(np.float64),
                n=rng.integers(1, 5, k).astype(np.int32))


def test_session_roundtrip(ck, tmp_path):
    az = _pkg.load()
    hp = az.ResNetHP(5, 128, (3, 3), 32, 32, batch_norm_momentum=0.6)
    rng = np.random.default_rng(1)
    best, cur = rng.normal(size=1000).astype(np.float32), rng.normal(size=1000).astype(np.float32)
    exp = _samples(257)
    d = str(tmp_path / "sessions" / "connect-four")
    assert not ck.valid_session_dir(d)
    ck.save_env(d, "connect-four", "resnet", hp, best, cur, 43, 7, exp, itc=12)
    assert ck.valid_session_dir(d)
    env = ck.load_env(d)
    assert env["itc"] == 12 and env["bestnn"]["kind"] == "resnet" and env["bestnn"]["game"] == "connect-four"
    assert (env["bestnn"]["blob"] == best).all() and (env["curnn"]["blob"] == cur).all()
    assert env["bestnn"]["hyperparams"]["num_blocks"] == 5 and env["bestnn"]["hyperparams"]["conv_kernel_size"] == [3, 3]
    for k in ("states", "pi", "z", "t", "n"):
        assert (env["experience"][k] == exp[k]).all() and env["experience"][k].dtype == exp[k].dtype
    # files the reference also writes keep its format
    assert json.load(open(os.path.join(d, "iter.txt"))) == 12
    assert json.load(open(os.path.join(d, "netparams.json")))["num_filters"] == 128
    assert not any(f.endswith(".tmp") for f in os.listdir(d))


def test_corrupt_files_are_rejected(ck, tmp_path):
    p = str(tmp_path / "net.azb")
    ck.save_network(p, "resnet", "connect-four", dict(num_blocks=1), np.arange(10, dtype=np.float32))
    raw = open(p, "rb").read()
    open(p, "wb").write(raw[:-4])
    with pytest.raises(ValueError):
        ck.load_network(p)
    open(p, "wb").write(b"XXXX" + raw[4:])
    with pytest.raises(ValueError):
        ck.load_network(p)
    m = str(tmp_path / "mem.azs")
    ck.save_memory(m, "connect-four", 43, 7, _samples(5))
    open(m, "ab").write(b"\0")
    with pytest.raises(ValueError):
        ck.load_memory(m)
    with pytest.raises(ValueError):
        ck.load_memory(p)   # a network file is not a memory file
    with pytest.raises(FileNotFoundError):
        ck.load_env(str(tmp_path / "nowhere"))


def test_torch_model_blob_through_checkpoint(ck, tmp_path):
    """The learning step's blob survives the file format and loads back into the torch mirror."""
    import torch
    import alphazero_jl_b200.learning as lrn
    az = _pkg.load()
    hp = az.ResNetHP(1, 16, (3, 3), 4, 4, batch_norm_momentum=0.6)
    torch.manual_seed(0)
    net = lrn.ResNetTorch((7, 6, 3), 7, hp)
    p = str(tmp_path / "curnn.azb")
    ck.save_network(p, "resnet", "connect-four", hp, net.to_blob())
    got = ck.load_network(p)
    net2 = lrn.ResNetTorch((7, 6, 3), 7, az.ResNetHP(**{k: (tuple(v) if isinstance(v, list) else v) for k, v in got["hyperparams"].items()}))
    net2.load_blob(got["blob"])
    assert (net2.to_blob() == net.to_blob()).all()
