"""The reference's DQN network (Net/DQNNet.py:6-66), same architecture, parameter names and shapes (state_dicts interchange),
with its two script-rot defects repaired: the Mish activation exists (the reference assigns `self.mish` without defining it,
DQNNet.py:31) and the number of input planes is a parameter (the reference hard-codes 4 while DDQN.py feeds 3 planes and
DQN.py feeds 1).  Stays plain PyTorch (cuDNN / cuBLAS): the north star keeps the DQN unchanged."""
import torch
import torch.nn as nn
import torch.nn.functional as F

# (name, in_channels, out_channels, kernel, padding, stride); None = the constructor's in_planes
_CONVS = (("conv1", None, 32, 3, 1, 1), ("conv2", 32, 32, 3, 1, 1), ("conv3", 32, 32, 3, 1, 1), ("conv4", 32, 64, 3, 1, 1),
          ("conv5", 64, 64, 3, 1, 1), ("conv6", 64, 64, 3, 1, 1), ("conv7", 64, 64, 7, 3, 2))
_DENSE = (("fc1", 64 * 3 * 3, 256), ("fc2", 256, 128), ("actor1", 128, 64), ("actor2", 64, 4))


class Net(nn.Module):
    def __init__(self, in_planes=4, batch_size=64, gamma=0.9):
        super().__init__()
        for name, cin, cout, k, pad, stride in _CONVS:
            setattr(self, name, nn.Conv2d(in_planes if cin is None else cin, cout, k, padding=pad, stride=stride))
        self.pool = nn.AvgPool2d(kernel_size=3, padding=1, stride=2)
        for name, fin, fout in _DENSE:
            setattr(self, name, nn.Linear(fin, fout))
        self.dropout = nn.Dropout(p=0.2)
        self.activation = self.mish
        self.batch_size, self.gamma = batch_size, gamma  # DQN.py:263,278 read these off the model

    @staticmethod
    def mish(x):
        """x * tanh(softplus(x)) (Net/ACNet.py:56-57) as PyTorch's single fused kernel"""
        return F.mish(x)

    def _residual_pair(self, x, first, second):
        return self.activation(second(self.activation(first(x))) + x)

    def forward(self, x):
        x = x.to(self.conv1.weight.device)
        if not torch.is_autocast_enabled():
            x = x.to(self.conv1.weight.dtype)
        act = self.activation
        x = self._residual_pair(act(self.conv1(x)), self.conv2, self.conv3)      # 12x12, 32 channels
        x = self._residual_pair(act(self.conv4(x)), self.conv5, self.conv6)      # 12x12, 64 channels
        x = act(self.conv7(self.pool(x))).flatten(1)                               # 6x6 -> 3x3 -> 576
        x = self.dropout(act(self.fc1(x)))
        x = self.dropout(act(self.fc2(x)))
        return self.actor2(act(self.actor1(x)))

    def act(self, x, extra=None):
        return torch.argmax(self(x), dim=1)
