"""One small bf16 + one fp32 training step, for `compute-sanitizer --tool memcheck python tools/sanitize_step.py`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from carla_imitation_learning_b200 import FusedAdam, stage_frames, stage_gray, sliding_window
from carla_imitation_learning_b200.data import synthetic_sequence
from src.architectures.nets import ConvNet1
dev = torch.device("cuda", 0)
for mode, B in (("bf16", 5), ("bf16", 9), ("fp32", 3)):
    frames, labels = synthetic_sequence(1, B + 4)
    fr = torch.from_numpy(frames).to(dev)
    y = torch.from_numpy(labels[4:4 + B]).to(dev)
    torch.manual_seed(12345)
    net = ConvNet1({"obs_size": 4, "n_actions": 9, "precision": mode}).to(dev)
    eng = net.engine()
    opt = FusedAdam(list(net.parameters()), lr=1e-3)
    x = stage_frames(fr) if mode == "bf16" else sliding_window(stage_gray(fr))
    b = eng.train_forward_backward(x, y)
    opt.step_flat(eng.grads)
    if mode == "bf16":
        eng.pack_weights()
        net.act(x)
    torch.cuda.synchronize()
    print(mode, B, "loss", float(b.loss), "err flag", int(eng.err_flag.item()), flush=True)
print("sanitize_step done")
