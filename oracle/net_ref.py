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
        # an fp32 reference: no TF32 in cuDNN convolutions / cuBLAS matmuls when it runs on a GPU
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
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


class _RoundBF16(torch.autograd.Function):
    """y = bf16(x) in forward, dx = bf16(dy) in backward: a tensor the device path STORES as bf16 (activations and the
    gradients flowing through them)."""

    @staticmethod
    def forward(ctx, x):
        return x.bfloat16().float()

    @staticmethod
    def backward(ctx, g):
        return g.bfloat16().float()


class _RoundFwd(torch.autograd.Function):
    """bf16 operand image of an fp32 master tensor: rounded in forward, gradient passed through in fp32."""

    @staticmethod
    def forward(ctx, x):
        return x.bfloat16().float()

    @staticmethod
    def backward(ctx, g):
        return g


class RefTrainer:
    """fp32 PyTorch restatement of `Network::train_inner` + `Adam` for Net6 / Net5 (alpha-tak/src/model/network.rs:37-97,
    net6.rs:111-122, net5.rs:113-124): forward_training (BatchNorm on batch statistics, momentum 0.1, eps 1e-5 -- tch's BatchNormConfig
    defaults), loss = -sum(pi * log_softmax)/B + sum((z - v)^2)/B, gradients accumulated over chunks, then
    torch.optim.Adam(lr, weight_decay) -- the optimiser tch's nn::Adam {wd} builds.  TEST INFRASTRUCTURE (parity unpinned
    against the reference binary: libtorch 1.11 is not available here)."""

    def __init__(self, arch: int, blob: np.ndarray, device="cpu", lr=1e-4, wd=1e-4, emulate_bf16=False):
        """emulate_bf16: round exactly where the device path does -- conv weights to bf16 operand images, every stored
        activation (raw conv output, layer output) and the gradient flowing through it to bf16 -- keeping fp32
        accumulation, fp32 BatchNorm arithmetic and fp32 weight gradients.  It separates "the kernels compute what
        autograd computes" (tight tolerance against this mode) from "bf16 storage costs precision" (loose tolerance
        against plain fp32)."""
        assert arch in (5, 6)
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        self.emulate = emulate_bf16
        self.arch, self.blocks, self.device = arch, (16 if arch == 6 else 8), device
        self.names = [name for name, _ in W.spec(arch)]
        self.t = {}
        for k, v in W.split(np.array(blob, dtype=np.float32, copy=True), arch).items():
            x = torch.from_numpy(np.ascontiguousarray(v)).to(device)
            if not k.endswith("running_mean") and not k.endswith("running_var"):
                x.requires_grad_(True)
            self.t[k] = x
        self.params = [v for v in self.t.values() if v.requires_grad]
        self.opt = torch.optim.Adam(self.params, lr=lr, weight_decay=wd)
        self.opt.zero_grad()

    def _bn(self, x, p):
        t = self.t
        return F.batch_norm(x, t[p + ".running_mean"], t[p + ".running_var"], t[p + ".weight"], t[p + ".bias"],
                            training=True, momentum=0.1, eps=1e-5)

    def forward_training(self, x):
        t = self.t
        ra = _RoundBF16.apply if self.emulate else (lambda v: v)
        rw = _RoundFwd.apply if self.emulate else (lambda v: v)

        def conv(v, name):
            return ra(F.conv2d(v, rw(t[name + ".weight"]), t[name + ".bias"], padding=1))

        s = ra(F.relu(self._bn(conv(ra(x), "initial_conv"), "initial_bn")))
        for b in range(self.blocks):
            p = f"block{b}."
            y = ra(F.relu(self._bn(conv(s, p + "conv1"), p + "bn1")))
            y = self._bn(conv(y, p + "conv2"), p + "bn2")
            s = ra(F.relu(y + s))
        if self.arch == 6:
            logits = F.conv2d(s, rw(t["policy_conv.weight"]), t["policy_conv.bias"], padding=1).reshape(s.shape[0], -1)
        else:   # net5.rs:106-108
            logits = F.linear(s.reshape(s.shape[0], -1), rw(t["policy_fc.weight"]), t["policy_fc.bias"])
        logp = torch.log_softmax(logits, dim=1)
        value = torch.tanh(F.linear(s.reshape(s.shape[0], -1), t["value_fc.weight"], t["value_fc.bias"]))
        return logp, value

    def chunk(self, inputs: np.ndarray, pi: np.ndarray, z: np.ndarray):
        x = torch.from_numpy(inputs).to(self.device)
        p = torch.from_numpy(pi).to(self.device)
        zz = torch.from_numpy(z).to(self.device).unsqueeze(1)
        logp, value = self.forward_training(x)
        b = x.shape[0]
        loss_p = -(p * logp).sum() / b
        loss_z = (zz - value).square().sum() / b
        (loss_z + loss_p).backward()
        return float(loss_p), float(loss_z)

    def _blob_of(self, get):
        return np.concatenate([get(self.t[k]).detach().cpu().numpy().reshape(-1) for k in self.names]).astype(np.float32)

    def grads(self) -> np.ndarray:
        return self._blob_of(lambda v: v.grad if v.requires_grad and v.grad is not None else torch.zeros_like(v))

    def set_grads(self, blob: np.ndarray):
        for k, g in W.split(np.asarray(blob, dtype=np.float32), self.arch).items():
            if self.t[k].requires_grad:
                self.t[k].grad = torch.from_numpy(np.ascontiguousarray(g)).to(self.device).clone()

    def step(self):
        self.opt.step()
        self.opt.zero_grad()

    def blob(self) -> np.ndarray:
        return self._blob_of(lambda v: v)
