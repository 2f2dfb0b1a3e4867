"""Plain PyTorch policy / critic networks with the shapes franQ uses (franQ/Agent/models/{mlp,gaussian_mlp}.py): a
feed-forward net whose head sees the input and every hidden layer, an ensemble that concatenates its members' outputs
on the last dim (critic atoms = num_critics x num_q_predictions), and a tanh-squashed Gaussian policy.
These stay ordinary torch modules on purpose (BASELINE.json north_star); nothing here is on the CUDA hot path."""
import torch
from torch import nn


class SkipHeadMLP(nn.Module):
    def __init__(self, in_features, out_features, hidden_sizes):
        super().__init__()
        dims = [in_features] + list(hidden_sizes)
        self.hidden = nn.ModuleList([nn.Linear(a, b) for a, b in zip(dims[:-1], dims[1:])])
        self.act = nn.LeakyReLU()
        self.head = nn.Linear(sum(dims), out_features)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                nn.init.zeros_(m.bias)

    def forward(self, x):
        feats = [x]
        for lin in self.hidden:
            x = self.act(lin(x))
            feats.append(x)
        return self.head(torch.cat(feats, dim=-1))


class MLPEnsemble(nn.Module):
    def __init__(self, in_features, out_features, hidden_sizes, ensemble_size):
        super().__init__()
        self.nets = nn.ModuleList([SkipHeadMLP(in_features, out_features, hidden_sizes) for _ in range(ensemble_size)])

    def forward(self, x):
        return torch.cat([net(x) for net in self.nets], dim=-1)


class GaussianPolicy(SkipHeadMLP):
    """returns (action, log_prob [..., 1], tanh(mean)) like GaussianMLP.forward (gaussian_mlp.py:15-39)"""

    def __init__(self, in_features, action_dim, hidden_sizes, log_sig_min=-20.0, log_sig_max=2.0, epsilon=1e-4):
        super().__init__(in_features, 2 * action_dim, hidden_sizes)
        self.log_sig_min, self.log_sig_max, self.epsilon = log_sig_min, log_sig_max, epsilon

    def forward(self, state):
        mean, log_std = torch.chunk(super().forward(state), 2, dim=-1)
        log_std = log_std.clamp(self.log_sig_min, self.log_sig_max)
        # reparameterised sample and its log-density written out (torch.distributions validates its arguments with a
        # host-side .all(), which would put a sync in every step and cannot be captured in a CUDA graph)
        eps = torch.randn_like(mean)
        x = mean + log_std.exp() * eps
        action = torch.tanh(x)
        normal_log_prob = -0.5 * eps.pow(2) - log_std - 0.9189385332046727  # log(sqrt(2*pi))
        log_prob = (normal_log_prob - torch.log((1 - action.pow(2)) + self.epsilon)).sum(-1, keepdim=True)
        return action, log_prob, torch.tanh(mean)
