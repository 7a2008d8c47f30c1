"""Reward head p(r_t | h_t, s_t) — drop-in for the reference's ``utils/models/reward_model.py``.

On the default training path its loss is zeroed (predict_reward: False, reference
base/algo.py:200-201) so the algorithm layer skips its forward entirely; the module exists for the
parameter list / checkpoint layout and runs through the same MLP kernels when called."""
import torch
from torch import nn

from mrssm_b200 import ops
from utils.models.encoder import act_code


class RewardModel(nn.Module):
    def __init__(self, h_size, s_size, hidden_size, activation="relu"):
        super().__init__()
        self.activation = activation
        self.fc1 = nn.Linear(s_size + h_size, hidden_size)
        self.fc2 = nn.Linear(hidden_size, hidden_size)
        self.fc3 = nn.Linear(hidden_size, 1)
        self.modules = [self.fc1, self.fc2, self.fc3]

    def forward(self, h_t, s_t):
        T, B = h_t.shape[:2]
        y = ops.mlp(act_code(self.activation), False, 2, h_t.reshape(T * B, -1), s_t.reshape(T * B, -1),
                            self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias,
                            self.fc3.weight, self.fc3.bias)
        reward = y.reshape(T, B)
        return {"loc": reward, "scale": torch.ones_like(reward)}

    def get_log_prob(self, h_t, s_t, r_t):
        loc = self.forward(h_t, s_t)["loc"]
        return -0.5 * (r_t - loc) ** 2 - 0.9189385332046727      # log N(r; loc, 1)
