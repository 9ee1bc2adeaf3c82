import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from carla_imitation_learning_b200 import stage_gray, FusedAdam
from oracle import bc_oracle as O
from src.architectures.nets import ConvNet1
dev = torch.device("cuda", 0)
B, steps = 8, 450
frames, labels = O.synth_frames(11, 1000 * B + 4)
frames, labels = frames[:steps * B + 4], labels[:steps * B + 4]
lab = torch.from_numpy(labels).to(dev)
fr = torch.from_numpy(frames).to(dev)
def run(name, staged, fwd_mode, bwd_mode, layers_tc=None):
    torch.manual_seed(12345)
    net = ConvNet1({"obs_size": 4, "n_actions": 9}).to(dev)
    eng = net.engine()
    opt = FusedAdam(list(net.parameters()), lr=1e-3)
    gray = stage_gray(fr, dtype=staged)
    L = []
    for s in range(steps):
        x = gray.as_strided((B, 4, 256, 256), (65536, 65536, 256, 1), s * B * 65536)
        y = lab[s * B + 4: s * B + 4 + B]
        eng.conv_mode = fwd_mode
        if fwd_mode or bwd_mode:
            eng.pack_weights()
        b = eng.forward(x, y, backward=True)
        eng.conv_mode = bwd_mode
        if bwd_mode and not b.act_bf16:
            raise SystemExit("need bf16 acts")
        eng.backward(b)
        opt.step_flat(eng.grads)
        L.append(float(b.loss))
    eng.check_device_errors()
    print(f"{name:28s}", " ".join(f"{np.mean(L[i:i+30]):.3f}" for i in range(0, steps, 30)))
run("fp32 all", torch.float32, 0, 0)
run("bf16 input, fp32 kernels", torch.bfloat16, 0, 0)
run("tc fwd, fp32 bwd", torch.bfloat16, 1, 0)
run("tc fwd + dgrad", torch.bfloat16, 1, 2)
run("tc fwd + wgrad2-4", torch.bfloat16, 1, 4)
run("tc fwd + wgrad1", torch.bfloat16, 1, 8)
run("tc fwd, tc bwd", torch.bfloat16, 1, 15)
