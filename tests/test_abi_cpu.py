"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/azb200.h
declares, answers the pure-host game queries, and fails loudly (no CPU fallback) without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import _pkg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def az():
    _pkg.build()
    return _pkg.load()


def test_library_exports_every_declared_symbol(az):
    hdr = open(os.path.join(ROOT, "include", "azb200.h")).read()
    declared = set(re.findall(r"\b(az_[a-z_0-9]+)\s*\(", hdr))
    declared -= {"az_status"}
    assert declared == set(az.ABI_SYMBOLS), declared ^ set(az.ABI_SYMBOLS)
    L = az.lib()
    for name in sorted(declared):
        assert hasattr(L, name), name
    assert L.az_version() == 1


def test_game_queries_match_oracle(az, oz):
    """The product's host game helpers (same inline code as the kernels) against the CPU oracle on random playouts."""
    rng = np.random.default_rng(1)
    for name in ["connect-four", "tictactoe", "mancala"]:
        gs = az.GameSpec(name)
        gid = oz.game_id(name)
        assert gs.num_actions == oz.num_actions(gid) and gs.state_bytes == oz.state_bytes(gid)
        assert gs.state_dim == oz.state_dim(gid)
        for _ in range(40):
            g = oz.GameEnv(gid)
            s = gs.init_state()
            assert bytes(s) == g.state()
            while not g.terminated():
                m = g.actions_mask()
                assert (gs.actions_mask(s) == m).all()
                assert (gs.vectorize_state(s) == oz.vectorize_state(gid, g.state())).all()
                assert gs.heuristic_value(s) == oz.lib().oz_heuristic_value(C.byref(g.g))      # GI.heuristic_value
                a = int(rng.choice(np.flatnonzero(m)))
                g.play(a)
                s, term, wr = gs.play(s, a)
                assert bytes(s) == g.state() and term == g.terminated() and wr == g.white_reward()


def test_random_positions_match_oracle(az, oz):
    gs = az.GameSpec("connect-four")
    a = gs.random_positions(0xA17A2E80, 64, 30)
    b = oz.random_positions(oz.game_id("connect-four"), 0xA17A2E80, 64, 30)
    assert (a == b).all()


def test_no_cpu_fallback(az):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(az.AzError) as e:
        az.Context(0)
    assert "no CPU fallback" in str(e.value)


def test_grid_world_host_helpers_match_oracle(az, oz):
    gs = az.GameSpec("grid-world")
    gid = oz.game_id("grid-world")
    assert gs.num_actions == 4 and gs.state_bytes == 2 and gs.state_dim == (10, 10, 1) and gs.max_plies == 201
    a = gs.random_positions(77, 50)
    b = oz.random_positions(gid, 77, 50)
    assert (a == b).all()
    for s in a:
        assert (gs.vectorize_state(s) == oz.vectorize_state(gid, bytes(s))).all()
        for act in range(4):                     # az_game_play is the noise-free step
            g = oz.GameEnv(gid, bytes(s))
            g.play(act, [0.9, 0.0])
            ns, term, wr = gs.play(s, act)
            assert bytes(ns) == g.state() and term == g.terminated() and wr == g.white_reward()


def test_plain_c_host_compiles_links_and_fails_loudly(az, tmp_path):
    """include/azb200.h is valid C99 (no C++ / torch types), a plain-C host links against the library, and without a GPU the
    first call fails with AZ_ECUDA and a message instead of falling back to a CPU path (examples/selfplay_c_abi.c)."""
    import shutil
    import subprocess
    cc = shutil.which("gcc")
    if cc is None:
        pytest.skip("gcc not available")
    libdir = os.path.join(ROOT, "alphazero.jl_b200")
    exe = str(tmp_path / "selfplay_c")
    subprocess.check_call([cc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "selfplay_c_abi.c"), "-L" + libdir, "-lazb200", "-Wl,-rpath," + libdir, "-o", exe])
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    r = subprocess.run([exe, "8", "16"], capture_output=True, text=True, timeout=120)
    if has_gpu:
        assert r.returncode == 0 and "games 8" in r.stdout, r.stderr
    else:
        assert r.returncode == 3 and "no CUDA device" in r.stderr and "no CPU fallback" in r.stderr


def test_evaluation_helpers_host_logic(az, monkeypatch):
    """compare_networks / TernaryOutcomeStatistics (src/training.jl:157-174, src/benchmark.jl:104-121): host composition only,
    with the engine calls stubbed (no GPU here)."""
    calls = []

    def fake_simulate(ctx, gs, net, params, seed=0, game_simulated=None, first_game_index=0, baseline=None, gamma=None):
        calls.append((net, baseline, gamma))
        r = np.array([1.0, 0.0, -1.0, 1.0]) if net == "new" else np.array([0.0, 0.0, -1.0, 1.0])
        return dict(game_rewards=r, redundancy=0.25 if net == "new" else 0.75)

    monkeypatch.setattr(az, "simulate", fake_simulate)

    class P:
        class mcts:
            gamma = 0.9
    ev = az.compare_networks(None, None, "new", "old", P)
    assert calls == [("new", "old", 0.9)] and ev.avgr == 0.25 and ev.redundancy == 0.25 and ev.baseline_rewards is None
    st = az.TernaryOutcomeStatistics(ev)
    assert (st.num_won, st.num_draw, st.num_lost) == (2, 1, 1)
    calls.clear()
    ev = az.compare_networks(None, None, "new", "old", P, two_players=False)
    assert [c[:2] for c in calls] == [("new", None), ("old", None)] and ev.avgr == 0.25 - 0.0 and ev.redundancy == 0.5
    assert (ev.baseline_rewards == [0.0, 0.0, -1.0, 1.0]).all()


def test_fresh_network_initialisers_match_the_restatement(az):
    """`fresh_resnet_blob` / `fresh_simplenet_blob` (what scripts and benches load, no oracle import on the product side) build
    the blob of a freshly constructed Flux model in the order `az_net_load` expects: same length and same values as the
    oracle-side builder with the same seed."""
    from oracle import netref
    for game, hp, hd in [("connect-four", az.ResNetHP(5, 128, (3, 3), 32, 32),
                          dict(num_blocks=5, num_filters=128, conv_kernel_size=(3, 3), num_policy_head_filters=32, num_value_head_filters=32)),
                         ("mancala", az.ResNetHP(2, 64, (3, 3), 2, 1),
                          dict(num_blocks=2, num_filters=64, conv_kernel_size=(3, 3), num_policy_head_filters=2, num_value_head_filters=1))]:
        gs = az.GameSpec(game)
        b = az.fresh_resnet_blob(gs, hp, seed=3)
        assert len(b) == netref.num_params(gs.state_dim, gs.num_actions, hd)
        assert np.array_equal(b, netref.make_blob(gs.state_dim, gs.num_actions, hd, seed=3, randomize=False))
    gs = az.GameSpec("grid-world")
    for hp, hd in [(az.SimpleNetHP(100, 4), dict(width=100, depth_common=4, use_batch_norm=False)),
                   (az.SimpleNetHP(64, 2, 2, 0, True), dict(width=64, depth_common=2, depth_phead=2, depth_vhead=0, use_batch_norm=True))]:
        assert np.array_equal(az.fresh_simplenet_blob(gs, hp, seed=2), netref.simplenet_make_blob(gs.state_dim, 4, hd, seed=2, randomize=False))


def test_struct_layouts_of_header_and_ctypes_mirror_agree(az, tmp_path):
    """sizeof / offsetof of every parameter struct in include/azb200.h, as gcc lays them out, against the ctypes mirrors of the
    host package (field names taken from the header: a renamed, reordered or retyped field shows up here without a GPU)."""
    import shutil
    import subprocess
    cc = shutil.which("gcc")
    if cc is None:
        pytest.skip("gcc not available")
    hdr = open(os.path.join(ROOT, "include", "azb200.h")).read()
    mirrors = {"az_mcts_params": az._MctsParams, "az_minmax_params": az._MinMaxParams, "az_sim_params": az._SimParams,
               "az_resnet_hp": az._ResNetHP, "az_simplenet_hp": az._SimpleNetHP}
    structs = {}
    for body, name in re.findall(r"typedef struct \{(.*?)\} (az_\w+);", hdr, flags=re.S):
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        structs[name] = [re.sub(r"\[.*", "", d.strip().split()[-1]) for d in body.split(";") if d.strip()]
    assert set(structs) == set(mirrors), set(structs) ^ set(mirrors)
    src = ["#include <stdio.h>", "#include <stddef.h>", '#include "azb200.h"', "int main(void) {"]
    for name, fields in structs.items():
        src.append('printf("%s %%zu\\n", sizeof(%s));' % (name, name))
        for f in fields:
            src.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (name, f, name, f))
    src.append("return 0; }")
    cfile, exe = tmp_path / "layout.c", tmp_path / "layout"
    cfile.write_text("\n".join(src))
    subprocess.check_call([cc, "-std=c99", "-I" + os.path.join(ROOT, "include"), str(cfile), "-o", str(exe)])
    got = dict(line.split() for line in subprocess.check_output([str(exe)], text=True).splitlines())
    for name, fields in structs.items():
        m = mirrors[name]
        assert [f for f, _ in m._fields_] == fields, name
        assert int(got[name]) == C.sizeof(m), name
        for f in fields:
            assert int(got[name + "." + f]) == getattr(m, f).offset, (name, f)
