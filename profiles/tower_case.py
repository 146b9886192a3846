"""The hand-written tower in isolation (for ncu): a few forward passes at the self-play batch."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from harmonies_alphazero_b200 import net as hnet  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
torch.manual_seed(0)
model = hnet.AlphaZeroNet.from_config(hnet.DEFAULT_MODEL_CONFIG).eval()
net = hnet.InferenceNet(model, tower="hand")
x0 = net.hand.x0_buffer(B)
board = torch.zeros((B, 40, 5, 7), dtype=torch.bfloat16, device="cuda").contiguous(memory_format=torch.channels_last)
board[:, :38] = (torch.rand((B, 38, 5, 7), device="cuda") < 0.15).to(torch.bfloat16)
net.hand.to_tiles(board, 40, True, x0)
glob = torch.rand((B, 42), device="cuda").to(torch.bfloat16)
for _ in range(4):
    net.forward_tiles(x0, glob, B)
torch.cuda.synchronize()
print("ok")
