"""The network at its DEFAULT size (config.py:18-29, the network every throughput number uses) against
the reference's own AlphaZeroModel (model.py:277-357): tests/golden/net_default.npz holds inputs and the
reference's outputs for weights that both sides rebuild from oracle/synth_weights (numpy PCG64).
CPU part: this package's AlphaZeroNet in fp32.  GPU part: the folded bf16 inference network with the
hand-written sm_100a tower and with cuDNN."""

import numpy as np
import pytest
import torch

from tests.conftest import load_golden


def _model():
    from harmonies_alphazero_b200 import net as hnet
    from oracle import synth_weights

    g = load_golden("net_default")
    model = hnet.AlphaZeroNet.from_config(hnet.DEFAULT_MODEL_CONFIG)
    synth_weights.fill_(model, int(g["seed"]))
    return g, model.eval()


def test_default_size_network_reproduces_the_reference_fp32():
    g, model = _model()
    with torch.no_grad():
        logits, value = model(torch.from_numpy(g["board"]), torch.from_numpy(g["glob"]))
    assert np.abs(logits.numpy() - g["logits"]).max() <= 2e-5          # fp32, same op order up to conv algorithm choice
    assert np.abs(value.numpy().reshape(-1) - g["value"]).max() <= 2e-5
    assert np.abs(torch.softmax(logits, dim=1).numpy() - g["probs"]).max() <= 1e-6


@pytest.mark.gpu
def test_hand_written_tower_network_against_the_reference_model():
    """bf16 tolerance: 17 convolution layers of bf16 activations; the hand-written tower must be at
    least as close to the reference's fp32 outputs as the cuDNN tower is (factor 2 slack) and within
    3 % of the output scale in absolute terms."""
    from harmonies_alphazero_b200 import net as hnet

    g, model = _model()
    B = g["board"].shape[0]
    b40 = torch.zeros((B, 40, 5, 7), dtype=torch.bfloat16, device="cuda").contiguous(memory_format=torch.channels_last)
    b40[:, :38] = torch.from_numpy(g["board"]).cuda().to(torch.bfloat16)
    gl = torch.from_numpy(g["glob"]).cuda().to(torch.bfloat16)
    err = {}
    for tower in ("hand", "cudnn"):
        inf = hnet.InferenceNet(model, device="cuda", tower=tower)
        logits, value = inf(b40, gl)
        x = inf.tower_out(b40)[:4].float().cpu().numpy()
        err[tower] = (np.abs(logits.cpu().numpy() - g["logits"]).max(), np.abs(value.cpu().numpy() - g["value"]).max(),
                      np.abs(x - g["tower_sample"].astype(np.float32)).max())
        if tower == "hand":
            assert inf.wants_tiles
            x0 = inf.hand.x0_buffer(B)
            inf.hand.to_tiles(b40, 40, True, x0)
            l2, v2 = inf.forward_tiles(x0, gl, B)
            assert np.abs(l2.cpu().numpy() - g["logits"]).max() <= max(2.0 * err[tower][0], 1e-3)
            probs = torch.softmax(l2, dim=1).cpu().numpy()
            assert np.abs(probs - g["probs"]).max() <= 0.03 * g["probs"].max()
    scale = float(np.abs(g["logits"]).max())
    assert err["hand"][0] <= max(2.0 * err["cudnn"][0], 0.03 * scale), err
    assert err["hand"][1] <= max(2.0 * err["cudnn"][1], 0.03), err
    assert err["hand"][2] <= max(2.0 * err["cudnn"][2], 0.03 * float(g["tower_abs_max"])), err
