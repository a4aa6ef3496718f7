"""The reference's DQN network (Net/DQNNet.py:6-66) with its two script-rot defects repaired in our own copy:
`mish` is defined (the reference assigns self.mish without defining it, DQNNet.py:31) and the number of input planes is a
parameter (the reference hard-codes 4 while DDQN.py feeds 3 and DQN.py feeds 1).  Architecture, layer names and
parameter shapes are otherwise identical, so state_dicts interchange; it stays plain PyTorch (cuDNN/cuBLAS)."""
import torch
import torch.nn as nn
import torch.nn.functional as F


class Net(nn.Module):
    def __init__(self, in_planes=4, batch_size=64, gamma=0.9):
        super().__init__()
        self.conv1 = nn.Conv2d(in_planes, 32, 3, padding=1)
        self.conv2 = nn.Conv2d(32, 32, 3, padding=1)
        self.conv3 = nn.Conv2d(32, 32, 3, padding=1)
        self.conv4 = nn.Conv2d(32, 64, 3, padding=1)
        self.conv5 = nn.Conv2d(64, 64, 3, padding=1)
        self.conv6 = nn.Conv2d(64, 64, 3, padding=1)
        self.pool = nn.AvgPool2d(kernel_size=3, padding=1, stride=2)
        self.conv7 = nn.Conv2d(64, 64, 7, padding=3, stride=2)
        self.fc1 = nn.Linear(64 * 3 * 3, 256)
        self.fc2 = nn.Linear(256, 128)
        self.actor1 = nn.Linear(128, 64)
        self.actor2 = nn.Linear(64, 4)
        self.dropout = nn.Dropout(p=0.2)
        self.activation = self.mish
        self.batch_size, self.gamma = batch_size, gamma  # DQN.py:263,278 read these off the model

    @staticmethod
    def mish(x):
        return F.mish(x)  # == x * tanh(softplus(x)) (Net/ACNet.py:56-57), as one fused PyTorch kernel instead of three

    def forward(self, x):
        x = x.to(self.conv1.weight.device)
        if not torch.is_autocast_enabled():
            x = x.to(self.conv1.weight.dtype)
        x = self.activation(self.conv1(x))
        skip = x
        x = self.activation(self.conv2(x))
        x = self.activation(self.conv3(x) + skip)
        x = self.activation(self.conv4(x))
        skip = x
        x = self.activation(self.conv5(x))
        x = self.activation(self.conv6(x) + skip)
        x = self.pool(x)
        x = self.activation(self.conv7(x))
        x = x.reshape(-1, 64 * 3 * 3)
        x = self.dropout(self.activation(self.fc1(x)))
        x = self.dropout(self.activation(self.fc2(x)))
        return self.actor2(self.activation(self.actor1(x)))

    def act(self, x, extra=None):
        return torch.argmax(self(x), dim=1)
