"""Engine-side analogue of the reference's `dummy_run` (src/scripts/dummy_run.jl, test/runtests.jl:19-26): two tiny
AlphaZero iterations on Connect Four that touch every stage this repository implements -- self-play on the device engine,
device-side sample preparation, the torch learning step, an arena duel of the updated network against the previous best,
and a checkpoint -- to surface runtime errors before a long run.  Needs a B200 (`gpurun -- python scripts/dummy_run.py`).
Not part of the test suite and not a training loop: there is no UI, report or session management here."""
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _pkg  # noqa: E402


def main(num_iters=2, games=8, workers=8, nsims=16, blocks=1):
    az = _pkg.load()
    import alphazero_jl_b200.checkpoint as ck
    import alphazero_jl_b200.learning as lrn
    ctx = az.Context(0)
    gs = az.GameSpec("connect-four")
    hp = az.ResNetHP(blocks, 128, (3, 3), 32, 32, batch_norm_momentum=0.6)
    model = lrn.ResNetTorch(gs.state_dim, gs.num_actions, hp)
    best_blob = model.to_blob()
    bestnn = az.ResNet(ctx, gs, hp).load(best_blob)
    self_play = az.SelfPlayParams(az.MctsParams(cpuct=2.0, num_iters_per_turn=nsims, temperature=az.PLSchedule([0, 20, 30], [1.0, 1.0, 0.3]),
                                                dirichlet_noise_eps=0.25, dirichlet_noise_alpha=1.0),
                                  az.SimParams(num_games=games, num_workers=workers, batch_size=workers, reset_every=2))
    arena = az.SelfPlayParams(az.MctsParams(cpuct=2.0, num_iters_per_turn=nsims, temperature=az.ConstSchedule(0.2), dirichlet_noise_eps=0.05,
                                            dirichlet_noise_alpha=1.0),
                              az.SimParams(num_games=games, num_workers=workers, batch_size=workers, reset_every=2, flip_probability=0.5,
                                           alternate_colors=True))
    learning = lrn.LearningParams(lrn.Adam(2e-3), l2_regularization=1e-4, batch_size=64, loss_computation_batch_size=128,
                                  max_batches_per_checkpoint=2, num_checkpoints=1)
    memory = None
    for itc in range(num_iters):
        t0 = time.perf_counter()
        sp = az.SelfPlay(ctx, gs, bestnn, self_play, seed=1000 + itc)                 # self_play_step! (src/training.jl:275-300)
        sp.start(games, 0)
        sp.wait()
        new = az.Samples.from_selfplay(sp)
        sp.close()
        if memory is None:
            memory = new
        else:
            memory, old = memory.concat(new), memory
            old.close(); new.close()
        aug = memory.augment_with_symmetries()                                        # learning_step! (src/training.jl:199-202)
        mrg = aug.merge_by_state()                                                    # Trainer (src/learning.jl:104-108)
        data = mrg.convert(az.LOG_WEIGHT)
        aug.close(); mrg.close()
        tr = lrn.Trainer(model, data, learning, device="cuda", seed=itc)
        losses = tr.batch_updates(min(learning.max_batches_per_checkpoint, max(1, tr.num_batches_total())))
        cur_blob = tr.get_trained_network_blob()
        curnn = az.ResNet(ctx, gs, hp).load(cur_blob)
        ev = az.compare_networks(ctx, gs, curnn, bestnn, arena, seed=2000 + itc)      # src/training.jl:157-174
        print("iteration %d: %d samples in memory, %d distinct positions, loss %s, arena avg reward %+.2f (redundancy %.2f), %.1f s"
              % (itc + 1, len(memory), len(data["W"]), ["%.3f" % x for x in losses], ev.avgr, ev.redundancy, time.perf_counter() - t0))
        if ev.avgr >= 0.0:                                                            # arena.update_threshold (src/training.jl:219-233)
            bestnn.close()
            bestnn, best_blob = curnn, cur_blob
        else:
            curnn.close()
    d = tempfile.mkdtemp(prefix="azb200-session-")
    ck.save_env(d, "connect-four", "resnet", hp, best_blob, model.to_blob(), gs.state_bytes, gs.num_actions, memory.fetch(), itc=num_iters)
    env = ck.load_env(d)
    assert env["itc"] == num_iters and len(env["experience"]["z"]) == len(memory) and (env["bestnn"]["blob"] == np.asarray(best_blob)).all()
    print("checkpoint written to", d)
    memory.close()
    bestnn.close()
    ctx.close()


if __name__ == "__main__":
    main()
