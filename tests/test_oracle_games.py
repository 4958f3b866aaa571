"""Pins the oracle's game rules against the reference's own known-answer data and
restates the GameInterface invariants of src/scripts/test_game.jl:37-110."""
import ctypes as C
import os

import numpy as np
import pytest

PONS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pons")
FILES = ["Test_L1_R1", "Test_L1_R2", "Test_L1_R3", "Test_L2_R1", "Test_L2_R2", "Test_L3_R1"]


def _lines(name):
    with open(os.path.join(PONS, name)) as f:
        for ln in f:
            mv, sc = ln.split()
            yield mv, int(sc)


@pytest.mark.parametrize("name", FILES)
def test_c4_pons_positions_are_legal_and_nonterminal(oz, name):
    """scripts/pons_benchmark.jl:92-99 replays the digits with GI.play!: every move must be
    legal and every position reached must be non-terminal (the solver only scores those)."""
    gid = oz.game_id("connect-four")
    n = 0
    for mv, _ in _lines(name):
        g = oz.GameEnv(gid)
        for ch in mv:
            a = int(ch) - 1
            assert not g.terminated()
            assert g.actions_mask()[a]
            g.play(a)
        assert not g.terminated()
        # set_state!/current_state round trip (test_game.jl:53-60)
        g2 = oz.GameEnv(gid, g.state())
        assert g2.state() == g.state() and not g2.terminated()
        assert g.white_playing() == (len(mv) % 2 == 0)
        n += 1
    assert n == 1000


def test_c4_endgame_scores_match_solver(oz):
    """Exhaustive negamax over the oracle's rules reproduces the exact solver score on the
    1000 end-game positions of Test_L3_R1 (pins win detection, legality and draw handling)."""
    gid = oz.game_id("connect-four")
    L = oz.lib()
    L.oz_c4_solve.argtypes = [C.c_void_p]
    for mv, sc in _lines("Test_L3_R1"):
        g = oz.GameEnv(gid)
        for ch in mv:
            g.play(int(ch) - 1)
        s = np.zeros(48, np.uint8)
        s[:43] = np.frombuffer(g.state(), np.uint8)
        assert L.oz_c4_solve(s.ctypes.data) == sc, mv


@pytest.mark.parametrize("name", ["connect-four", "tictactoe", "mancala"])
def test_game_interface_invariants(oz, name):
    """src/scripts/test_game.jl:37-110 on random playouts."""
    gid = oz.game_id(name)
    A = oz.num_actions(gid)
    rng = np.random.default_rng(0)
    for _ in range(50):
        g = oz.GameEnv(gid)
        assert g.white_playing()  # white plays first (test_game.jl:46)
        steps = 0
        while True:
            st = g.state()
            x = oz.vectorize_state(gid, st)
            assert x.dtype == np.float32 and x.shape == oz.state_dim(gid)
            m = g.actions_mask()
            assert m.shape == (A,) and m.dtype == bool
            if g.terminated():
                break
            assert m.any()  # terminal or legal moves exist (test_game.jl:82-85)
            c = g.clone()
            a = int(rng.choice(np.flatnonzero(m)))
            g.play(a)
            assert c.state() == st  # states are persistent after play! (test_game.jl:95-101)
            assert isinstance(g.white_reward(), float)
            g2 = oz.GameEnv(gid, g.state())
            assert g2.state() == g.state()
            assert g2.terminated() == g.terminated()
            if not g.terminated():
                assert (g2.actions_mask() == g.actions_mask()).all()
            steps += 1
            assert steps < 500


def test_c4_vectorize_layout(oz):
    """games/connect-four/game.jl:226-241: channels [empty, current player, opponent], index [col,row,c]."""
    gid = oz.game_id("connect-four")
    g = oz.GameEnv(gid)
    g.play(3)  # white in column 4, row 1
    x = oz.vectorize_state(gid, g.state())  # black to play -> colours flipped
    assert x[3, 0, 2] == 1 and x[3, 0, 1] == 0 and x[3, 0, 0] == 0
    assert x[:, :, 0].sum() == 41
    g.play(3)
    x = oz.vectorize_state(gid, g.state())  # white to play
    assert x[3, 0, 1] == 1 and x[3, 1, 2] == 1


def test_mancala_rules(oz):
    gid = oz.game_id("mancala")
    g = oz.GameEnv(gid)
    # 3 seeds from house 3 end in the store -> free turn (games/mancala/game.jl:173-174)
    g.play(2)
    assert g.white_playing()
    s = np.frombuffer(g.state(), np.uint8)
    assert s[0] == 1 and s[2 + 0 + 2 * 2] == 0 and s[2 + 0 + 2 * 1] == 4 and s[2 + 0 + 2 * 0] == 4
    # black-to-move states all vectorize like the initial board (flip_colors quirk, :224-229)
    g.play(0)
    assert not g.white_playing()
    x = oz.vectorize_state(gid, g.state())
    x0 = oz.vectorize_state(gid, oz.GameEnv(gid).state())
    assert (x == x0).all()


def test_tictactoe_win(oz):
    gid = oz.game_id("tictactoe")
    g = oz.GameEnv(gid)
    for a in [0, 3, 1, 4, 2]:
        assert not g.terminated()
        g.play(a)
    assert g.terminated() and g.white_reward() == 1.0


def test_grid_world_rules(oz):
    """games/grid-world/game.jl:14-59 through src/common_rl_intf.jl:118-160."""
    gid = oz.game_id("grid-world")
    g = oz.GameEnv(gid, bytes([5, 5]))
    assert g.white_playing() and g.actions_mask().all() and not g.terminated() and g.white_reward() == 0.0
    g.play(0, [0.9, 0.0])                      # no noise: action 0 = (+1, 0)
    assert g.state() == bytes([6, 5]) and g.white_reward() == 0.0
    g.play(0, [0.1, 0.60])                     # noise: random action floor(4*0.6) = 2 = (0, +1)
    assert g.state() == bytes([6, 6])
    g = oz.GameEnv(gid, bytes([10, 10]))
    g.play(0, [0.9, 0.0])                      # clamped at the border
    assert g.state() == bytes([10, 10])
    g = oz.GameEnv(gid, bytes([8, 3]))
    g.play(0, [0.9, 0.0])                      # (9,3): reward +10, terminal
    assert g.state() == bytes([9, 3]) and g.white_reward() == 10.0 and g.terminated()
    g = oz.GameEnv(gid, bytes([4, 5]))
    g.play(2, [0.9, 0.0])                      # (4,6): reward -5, terminal
    assert g.white_reward() == -5.0 and g.terminated()
    g = oz.GameEnv(gid, bytes([1, 1]))         # episode bound: terminal once time > 200 (game.jl:40-41), not part of the state
    for i in range(201):
        assert not g.terminated()
        g.play(1, [0.9, 0.0])
    assert g.terminated() and g.state() == bytes([1, 1])
    x = oz.vectorize_state(gid, bytes([3, 7]))
    assert x.shape == (10, 10, 1) and x[2, 6, 0] == 1 and x.sum() == 1


def test_oracle_struct_layouts_match_the_ctypes_mirrors(oz, tmp_path):
    """The checker's own boundary: oz_game / oz_mcts_params / oz_trace / oz_rollout_ctx as gcc lays them out against the
    ctypes mirrors of oracle/oracle.py (every field offset and the sizes)."""
    import ctypes as C
    import os
    import re
    import shutil
    import subprocess
    cc = shutil.which("gcc")
    if cc is None:
        pytest.skip("gcc not available")
    here = os.path.dirname(os.path.abspath(oz.__file__))
    mirrors = {"oz_game": oz.Game, "oz_mcts_params": oz.MctsParams, "oz_trace": oz.Trace, "oz_rollout_ctx": oz.RolloutCtx}
    src = ["#include <stdio.h>", "#include <stddef.h>", '#include "az_oracle.h"', "int main(void) {"]
    for name, m in mirrors.items():
        src.append('printf("%s %%zu\\n", sizeof(%s));' % (name, name))
        for f, _ in m._fields_:
            src.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (name, f, name, f))
    src.append("return 0; }")
    cfile, exe = tmp_path / "ozlayout.c", tmp_path / "ozlayout"
    cfile.write_text("\n".join(src))
    subprocess.check_call([cc, "-std=c11", "-I" + here, str(cfile), "-o", str(exe)])
    got = dict(line.split() for line in subprocess.check_output([str(exe)], text=True).splitlines())
    for name, m in mirrors.items():
        assert int(got[name]) == C.sizeof(m), name
        for f, _ in m._fields_:
            assert int(got[name + "." + f]) == getattr(m, f).offset, (name, f)
    # and the mirrors name every field of the C structs (count of declarators in the header)
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(here, "az_oracle.h")).read(), flags=re.S)
    for name, m in mirrors.items():
        body = re.search(r"typedef struct \{([^}]*)\} %s;" % name, hdr, flags=re.S).group(1)
        n_decl = sum(len(d.split(",")) for d in body.split(";") if d.strip())
        assert n_decl == len(m._fields_), (name, n_decl, len(m._fields_))
