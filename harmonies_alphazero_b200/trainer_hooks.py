"""Batch-level drop-in (boundary b2 of SURVEY.md §8): replacements for the bodies of
``Trainer.execute_self_play_phase`` (trainer.py:62-134) and ``self_play_worker``
(trainer.py:434-541) that run the games on the GPU core and hand the reference's
replay buffer exactly the tuples it expects.

    import trainer                                  # the reference module
    from harmonies_alphazero_b200 import trainer_hooks
    trainer_hooks.install(trainer)                  # Trainer now self-plays on the B200
"""

import random
import time

import torch

from .net import AlphaZeroNet, InferenceNet
from .selfplay import BatchedSelfPlay, SelfPlayConfig


def _inference_net(model, device, dtype):
    return InferenceNet(model, device=device, dtype=dtype)


def _fresh_seed():
    """The reference draws every game's tiles from the global ``random`` module
    (harmonies_engine.py:126), so consecutive phases never replay the same deals.  The batched
    engine derives a game's draw stream from (seed, game id): take the seed from the same global
    generator — fresh games every call, reproducible after ``random.seed(k)``."""
    return random.getrandbits(63)


def execute_self_play_phase(self, data_generating_manager, n_slots=4096, dtype=torch.bfloat16, device="cuda",
                            leaves_per_step=1, seed=None):
    """Bound as a method of the reference ``Trainer``.  Same contract as trainer.py:62-134:
    plays ``num_games_per_iter`` games with the data-generating (best) model and extends
    ``self.replay_buffer`` with (board, global, pi, z) CPU tensors of every completed game.
    ``leaves_per_step`` > 1 (a divisor of num_simulations) keeps that many simulations in flight
    per tree with virtual loss: not the reference's sequential search any more, but with few
    games per iteration (the reference's default is 25) it is what fills the network batch."""
    num_games = int(self.self_play_config["num_games_per_iter"])
    print(f"\n--- Starting Self-Play Phase ({num_games} games, batched on {device}) ---")
    t0 = time.time()
    model = data_generating_manager.model
    was_training = model.training
    model.eval()
    net = _inference_net(model, device, dtype)
    cfg = SelfPlayConfig.from_mcts_config(self.mcts_config, n_slots=max(1, min(n_slots, num_games)),
                                          leaves_per_step=int(leaves_per_step),
                                          seed=_fresh_seed() if seed is None else int(seed))
    traj = BatchedSelfPlay(net, cfg, device=device).play(num_games)
    # a deque(maxlen) keeps only the tail of an extend: build just those examples
    examples = traj.to_reference_examples(last=getattr(self.replay_buffer, "maxlen", None))
    self.replay_buffer.extend(examples)                      # trainer.py:127
    if was_training:
        model.train()
    print("--- Self-Play Finished ---")
    print(f"  Completed {traj.stats['games']}/{num_games} games.")
    print(f"  Added {len(examples)} examples.")
    print(f"  Buffer size: {len(self.replay_buffer)} / {self.replay_buffer.maxlen}")
    print(f"  Time taken: {time.time() - t0:.2f} seconds ({traj.stats['sims_per_s']:.0f} sims/s)")
    return traj.stats


def self_play_worker(args):
    """Signature-compatible with trainer.py:434: (state_dict, model_config, training_config,
    mcts_config, worker_device) -> list of (board, global, pi, z) for ONE game, [] on error."""
    model_state_dict, model_config, _training_config, mcts_config, worker_device = args
    try:
        model = AlphaZeroNet.from_config(model_config)
        model.load_state_dict(model_state_dict)
        model.eval()
        dev = "cuda" if str(worker_device) in ("cpu", "mps") else worker_device   # the engine is GPU-only
        net = _inference_net(model, dev, torch.float32)
        cfg = SelfPlayConfig.from_mcts_config(mcts_config, n_slots=1, use_cuda_graph=False, seed=_fresh_seed())
        return BatchedSelfPlay(net, cfg, device=dev).play(1).to_reference_examples()
    except Exception as e:  # noqa: BLE001  (the reference worker also returns [] on any failure, trainer.py:459-514)
        print(f"WORKER ERROR: {e}")
        return []


def evaluate_model(self, dtype=torch.bfloat16, device="cuda", eval_config=None, seed=None):
    """Bound as a method of the reference ``Trainer``: the candidate-vs-best match of
    trainer.py:293-431 with all ``eval_episodes`` games played concurrently (arena.play_match:
    candidate is player 0 in even games, every move a fresh search by the side to move's
    network under mcts_config_eval), followed by the reference's promotion rule: win rate over
    decided games (0.5 if none, :328-331) above ``eval_win_rate_threshold`` saves the candidate
    as the best model and reloads ``best_model_manager`` (:345-361).  Returns the match dict."""
    import sys

    from . import arena

    cfg = self.self_play_config
    n_games, threshold = int(cfg["eval_episodes"]), float(cfg["eval_win_rate_threshold"])
    if eval_config is None:                                  # trainer.py:392-394
        tm = sys.modules.get(type(self).__module__)
        name = "test_mcts_config_eval" if self.mcts_config.get("testing") else "mcts_config_eval"
        eval_config = getattr(tm, name)
    print(f"\n--- Starting Evaluation Phase ({n_games} games, batched on {device}) ---")
    t0 = time.time()
    nets = []
    for mgr in (self.model_manager, self.best_model_manager):
        was_training = mgr.model.training
        mgr.model.eval()
        nets.append(_inference_net(mgr.model, device, dtype))
        if was_training:
            mgr.model.train()
    res = arena.play_match(nets[0], nets[1], n_games, eval_config, device=device,
                           seed=_fresh_seed() if seed is None else int(seed))   # fresh deals per evaluation, like the reference
    decided = res["candidate_wins"] + res["best_wins"]
    win_rate = res["candidate_wins"] / decided if decided else 0.5
    res["win_rate"] = win_rate
    print(f"  Results: Candidate={res['candidate_wins']}, Best={res['best_wins']}, Draws/Errors={res['draws']}")
    print(f"  Candidate Win Rate (vs Best, excluding draws): {win_rate:.3f}")
    res["promoted"] = win_rate > threshold
    if res["promoted"]:
        print(f"  Candidate model passed threshold ({threshold:.2f}); updating '{self.best_model_filename}'.")
        self.model_manager.save_checkpoint(folder=cfg["checkpoint_folder"], filename=self.best_model_filename,
                                           iteration=cfg["num_iterations"])
        self.best_model_manager.load_checkpoint(folder=cfg["checkpoint_folder"], filename=self.best_model_filename)
    else:
        print(f"  Candidate model did not pass threshold ({threshold:.2f}). Best model remains unchanged.")
    print(f"  Time taken: {time.time() - t0:.2f} seconds")
    return res


def install(trainer_module, evaluation=True):
    """Monkey-patch the reference ``trainer`` module in place: self-play always, the arena
    evaluation (Trainer.evaluate_model) unless ``evaluation=False``."""
    trainer_module.Trainer.execute_self_play_phase = execute_self_play_phase
    trainer_module.self_play_worker = self_play_worker
    if evaluation:
        trainer_module.Trainer.evaluate_model = evaluate_model
    return trainer_module
