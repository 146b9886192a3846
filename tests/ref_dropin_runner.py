"""Runs the reference's OWN code (staged unmodified in baseline/_ref by baseline/stage_ref.py) on top of the
drop-in modules, in a process of its own (it aliases top-level module names):

    python tests/ref_dropin_runner.py unittest     the reference's tests/test_harmonies_engine.py
    python tests/ref_dropin_runner.py trainer      test_run.py's configuration through trainer.Trainer with
                                                   trainer_hooks.install(): self-play -> buffer.save_buffer ->
                                                   execute_training_phase -> checkpoint, then evaluate_model
Prints one JSON line; exit code 0 on success.
"""
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
sys.path.insert(0, ROOT)


def alias():
    from harmonies_alphazero_b200 import MCTS, harmonies_engine, process_game_state

    sys.modules["harmonies_engine"] = harmonies_engine
    sys.modules["process_game_state"] = process_game_state
    sys.modules["MCTS"] = MCTS
    sys.path.insert(0, REF)          # config, model, trainer, buffer, loggers, ... are the reference's own files


def run_unittest():
    import unittest

    alias()
    import importlib

    sys.path.insert(0, os.path.join(REF, "tests"))   # the reference's tests/ has no __init__.py
    suite = unittest.defaultTestLoader.loadTestsFromModule(importlib.import_module("test_harmonies_engine"))
    res = unittest.TextTestRunner(verbosity=0, stream=open(os.devnull, "w")).run(suite)
    out = {"ran": res.testsRun, "failures": [str(f[0]) + ": " + f[1][-300:] for f in res.failures],
           "errors": [str(e[0]) + ": " + e[1][-300:] for e in res.errors], "skipped": len(res.skipped)}
    print(json.dumps(out))
    return 0 if res.wasSuccessful() and res.testsRun >= 11 else 1


def run_trainer():
    alias()
    import torch

    import config
    import trainer
    from model import ModelManager

    from harmonies_alphazero_b200 import trainer_hooks

    trainer_hooks.install(trainer)
    work = tempfile.mkdtemp(prefix="hz_ref_trainer_")
    os.chdir(work)                                   # the test configs use relative folders
    sp = dict(config.test_self_play_config, num_games_per_iter=6, eval_episodes=4)
    mm = ModelManager(config.test_model_config, config.test_training_config)
    tr = trainer.Trainer(mm, config.test_mcts_config, sp, config.test_training_config)
    n0 = len(tr.replay_buffer)
    tr.run_training_loop()                           # test_run.py:19-22
    n1 = len(tr.replay_buffer)
    ex = tr.replay_buffer[-1]
    shapes = [tuple(t.shape) for t in ex]
    files = sorted(os.listdir(sp["checkpoint_folder"])) if os.path.isdir(sp["checkpoint_folder"]) else []
    buf_file = os.path.join(sp["replay_buffer_folder"], sp["replay_buffer_filename"])
    # a second self-play phase must play NEW games (fresh draw streams per phase)
    first_boards = torch.stack([e[0] for e in list(tr.replay_buffer)[:5]])
    tr.execute_self_play_phase(tr.best_model_manager)
    res = tr.evaluate_model()                        # trainer.py:293-366 through arena.play_match
    out = {"examples_before": n0, "examples_after_iteration": n1, "example_shapes": shapes,
           "checkpoint_files": files, "buffer_file_bytes": os.path.getsize(buf_file) if os.path.exists(buf_file) else 0,
           "buffer_len_after_second_phase": len(tr.replay_buffer), "first_boards_sum": float(first_boards.sum()),
           "eval": {k: res[k] for k in ("candidate_wins", "best_wins", "draws", "games", "promoted")}}
    print(json.dumps(out))
    ok = (n1 > n0 and shapes == [(38, 5, 7), (42,), (143,), (1,)] and out["buffer_file_bytes"] > 0 and files
          and res["games"] == 4 and res["candidate_wins"] + res["best_wins"] + res["draws"] == 4)
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(run_unittest() if sys.argv[1] == "unittest" else run_trainer())
