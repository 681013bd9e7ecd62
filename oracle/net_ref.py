"""TEST INFRASTRUCTURE -- NOT PRODUCT CODE.

fp32 PyTorch restatement of the reference's Net5 / Net6 `forward_mcts` (alpha-tak/src/model/net6.rs:70-109,
net5.rs:75-111, res_block.rs:13-23): conv3x3(pad 1) -> BN(eval, eps 1e-5) -> ReLU, residual blocks
conv-BN-ReLU-conv-BN-add-ReLU, policy = softmax over the WHOLE output vector, value = tanh(fc(flatten NCHW)).
Parity status: libtorch 1.11 / tch 0.7.2 are not available here and the reference ships no golden network
outputs, so network parity is "fp32 PyTorch restatement on seeded random weights" (parity unpinned against the
reference binary itself).  Used by tests/ and by bench.py's CPU legs only.
"""
import numpy as np
import torch
import torch.nn.functional as F

from tak_b200 import weights as W


class RefNet:
    def __init__(self, arch: int, blob: np.ndarray, device="cpu"):
        self.arch = arch
        self.n = arch
        self.blocks = 16 if arch == 6 else 8
        self.t = {k: torch.from_numpy(np.ascontiguousarray(v)).to(device) for k, v in W.split(blob, arch).items()}

    def _bn(self, x, p):
        t = self.t
        return F.batch_norm(x, t[p + ".running_mean"], t[p + ".running_var"], t[p + ".weight"], t[p + ".bias"],
                            training=False, eps=1e-5)

    @torch.no_grad()
    def forward_mcts(self, x: torch.Tensor):
        t = self.t
        s = F.relu(self._bn(F.conv2d(x, t["initial_conv.weight"], t["initial_conv.bias"], padding=1), "initial_bn"))
        for b in range(self.blocks):
            p = f"block{b}."
            y = F.relu(self._bn(F.conv2d(s, t[p + "conv1.weight"], t[p + "conv1.bias"], padding=1), p + "bn1"))
            y = self._bn(F.conv2d(y, t[p + "conv2.weight"], t[p + "conv2.bias"], padding=1), p + "bn2")
            s = F.relu(y + s)
        flat = s.reshape(s.shape[0], -1)
        if self.arch == 6:
            logits = F.conv2d(s, t["policy_conv.weight"], t["policy_conv.bias"], padding=1).reshape(s.shape[0], -1)
        else:
            logits = F.linear(flat, t["policy_fc.weight"], t["policy_fc.bias"])
        policy = torch.softmax(logits, dim=1)
        value = torch.tanh(F.linear(flat, t["value_fc.weight"], t["value_fc.bias"])).squeeze(1)
        return policy, value, logits
