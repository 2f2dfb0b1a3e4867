"""Plain PyTorch policy / critic networks with the shapes franQ uses (franQ/Agent/models/{mlp,gaussian_mlp}.py): a
feed-forward net whose head sees the input and every hidden layer, an ensemble that concatenates its members' outputs
on the last dim (critic atoms = num_critics x num_q_predictions), and a tanh-squashed Gaussian policy.
These stay ordinary torch modules on purpose (BASELINE.json north_star); nothing here is on the CUDA hot path."""
import torch
from torch import nn


class SkipHeadMLP(nn.Module):
    def __init__(self, in_features, out_features, hidden_sizes, activation_class=nn.LeakyReLU):
        super().__init__()
        dims = [in_features] + list(hidden_sizes)
        self.hidden = nn.ModuleList([nn.Linear(a, b) for a, b in zip(dims[:-1], dims[1:])])
        self.act = activation_class()
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

    @torch.no_grad()
    def load_reference_state_dict(self, sd, prefix=""):
        """Weights saved by franQ's SkipHeadMLP (models/mlp.py:62-90: `feature_extractor.{i}.0.{weight,bias}`, `head.{weight,bias}`)."""
        for i, lin in enumerate(self.hidden):
            lin.weight.copy_(torch.as_tensor(sd[f"{prefix}feature_extractor.{i}.0.weight"]))
            lin.bias.copy_(torch.as_tensor(sd[f"{prefix}feature_extractor.{i}.0.bias"]))
        self.head.weight.copy_(torch.as_tensor(sd[f"{prefix}head.weight"]))
        self.head.bias.copy_(torch.as_tensor(sd[f"{prefix}head.bias"]))

    def reference_state_dict(self, prefix=""):
        out = {}
        for i, lin in enumerate(self.hidden):
            out[f"{prefix}feature_extractor.{i}.0.weight"] = lin.weight.detach().clone()
            out[f"{prefix}feature_extractor.{i}.0.bias"] = lin.bias.detach().clone()
        out[f"{prefix}head.weight"], out[f"{prefix}head.bias"] = self.head.weight.detach().clone(), self.head.bias.detach().clone()
        return out


class FeedForwardEncoder(nn.Module):
    """franQ/Agent/components/encoder.py:13-58 in its feed-forward mode for 1-D observations: cat(obs_1d, achieved_goal,
    desired_goal) -> MLP(hidden_features) -> joiner MLP(out_features).  Ordinary torch (BASELINE.json north_star)."""

    def __init__(self, in_features, out_features, hidden_features=256, obs_1d_hidden_dims=(256,), joint_hidden_dims=(256,),
                 keys=("obs_1d", "achieved_goal", "desired_goal")):
        super().__init__()
        self.keys = tuple(keys)
        self.obs_1d = SkipHeadMLP(in_features, hidden_features, obs_1d_hidden_dims)
        self.joiner = SkipHeadMLP(hidden_features, out_features, joint_hidden_dims)

    def forward_train(self, xp):
        return self.joiner(self.obs_1d(torch.cat([xp[k] for k in self.keys if k in xp], dim=-1)))

    def load_reference_state_dict(self, sd, prefix=""):
        self.obs_1d.load_reference_state_dict(sd, prefix + "visible_layer_encoders.obs_1d.")
        self.joiner.load_reference_state_dict(sd, prefix + "joiner.")


class MLPEnsemble(nn.Module):
    def __init__(self, in_features, out_features, hidden_sizes, ensemble_size):
        super().__init__()
        self.nets = nn.ModuleList([SkipHeadMLP(in_features, out_features, hidden_sizes) for _ in range(ensemble_size)])

    def forward(self, x):
        return torch.cat([net(x) for net in self.nets], dim=-1)

    def load_reference_state_dict(self, sd, prefix=""):
        for e, net in enumerate(self.nets):
            net.load_reference_state_dict(sd, f"{prefix}nets.{e}.")


class BatchedMLPEnsemble(nn.Module):
    """The same function as MLPEnsemble (E independent SkipHeadMLPs, outputs concatenated on the last dim) with the members'
    weights stacked as [E, in, out] and every layer evaluated by one torch.baddbmm: E times fewer GEMM launches per forward and
    backward pass, which is what bounds a 4096-row learner step.  Still an ordinary torch module (no custom kernels).
    `load_member_state_dicts` / `member_state_dicts` convert from / to the per-member layout (franQ's `nets.{i}.hidden.{l}.weight`)."""

    def __init__(self, in_features, out_features, hidden_sizes, ensemble_size):
        super().__init__()
        self.E, self.in_features, self.out_features = int(ensemble_size), int(in_features), int(out_features)
        dims = [in_features] + list(hidden_sizes)
        self.dims = dims
        self.hidden_w = nn.ParameterList([nn.Parameter(torch.empty(self.E, a, b)) for a, b in zip(dims[:-1], dims[1:])])
        self.hidden_b = nn.ParameterList([nn.Parameter(torch.zeros(self.E, 1, b)) for b in dims[1:]])
        self.head_w = nn.Parameter(torch.empty(self.E, sum(dims), out_features))
        self.head_b = nn.Parameter(torch.zeros(self.E, 1, out_features))
        self.act = nn.LeakyReLU()
        for w in list(self.hidden_w) + [self.head_w]:
            for e in range(self.E):  # xavier_uniform_ per member on the [out, in] view, like nn.Linear in SkipHeadMLP
                nn.init.xavier_uniform_(w.data[e].t())

    def forward(self, x):
        lead, E, d = x.shape[:-1], self.E, self.dims
        x2 = x.reshape(-1, d[0])
        M = x2.shape[0]
        head = torch.split(self.head_w, d, dim=1)  # the head sees the input and every hidden layer: one weight block per source
        # the input is shared by the members: its two products (first hidden layer, input block of the head) are single GEMMs over
        # the members' weights laid side by side, read back as strided [E, M, .] views -- no expanded copy of x, no concatenation
        w0 = self.hidden_w[0].permute(1, 0, 2).reshape(d[0], E * d[1])
        h = self.act(torch.addmm(self.hidden_b[0].reshape(E * d[1]), x2, w0)).view(M, E, d[1]).transpose(0, 1)
        y = torch.addmm(self.head_b.reshape(E * self.out_features), x2, head[0].permute(1, 0, 2).reshape(d[0], E * self.out_features))
        y = torch.baddbmm(y.view(M, E, self.out_features).transpose(0, 1), h, head[1])
        for l in range(1, len(self.hidden_w)):
            h = self.act(torch.baddbmm(self.hidden_b[l], h, self.hidden_w[l]))
            y = torch.baddbmm(y, h, head[l + 1])
        return y.transpose(0, 1).reshape(*lead, E * self.out_features)

    @torch.no_grad()
    def load_member_state_dicts(self, members):
        """members[e] = state_dict of a SkipHeadMLP (hidden.{l}.weight [out, in], hidden.{l}.bias, head.weight, head.bias)."""
        for e, sd in enumerate(members):
            for l in range(len(self.hidden_w)):
                self.hidden_w[l][e].copy_(sd[f"hidden.{l}.weight"].t())
                self.hidden_b[l][e, 0].copy_(sd[f"hidden.{l}.bias"])
            self.head_w[e].copy_(sd["head.weight"].t())
            self.head_b[e, 0].copy_(sd["head.bias"])

    def load_reference_state_dict(self, sd, prefix=""):
        """Weights saved by franQ's MLPEnsemble (models/mlp.py:95-104: `nets.{e}.feature_extractor.{l}.0.weight`, `nets.{e}.head.weight`)."""
        members = []
        for e in range(self.E):
            m = {}
            for l in range(len(self.hidden_w)):
                m[f"hidden.{l}.weight"] = torch.as_tensor(sd[f"{prefix}nets.{e}.feature_extractor.{l}.0.weight"])
                m[f"hidden.{l}.bias"] = torch.as_tensor(sd[f"{prefix}nets.{e}.feature_extractor.{l}.0.bias"])
            m["head.weight"] = torch.as_tensor(sd[f"{prefix}nets.{e}.head.weight"])
            m["head.bias"] = torch.as_tensor(sd[f"{prefix}nets.{e}.head.bias"])
            members.append(m)
        self.load_member_state_dicts(members)

    def member_state_dicts(self):
        out = []
        for e in range(self.E):
            sd = {}
            for l in range(len(self.hidden_w)):
                sd[f"hidden.{l}.weight"] = self.hidden_w[l][e].t().detach().clone()
                sd[f"hidden.{l}.bias"] = self.hidden_b[l][e, 0].detach().clone()
            sd["head.weight"] = self.head_w[e].t().detach().clone()
            sd["head.bias"] = self.head_b[e, 0].detach().clone()
            out.append(sd)
        return out


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


class GumbelPolicy(SkipHeadMLP):
    """Discrete policy like franQ/Agent/models/gumbel_mlp.py:7-54: returns (straight-through one-hot action, log_prob [..., 1], logits).
    The relaxed sample is softmax((logits + Gumbel noise) / temperature); the action is its arg-max one-hot with the relaxed
    sample's gradient (:43-49); log_prob = sum(action * log_softmax(logits)) (:51-54).  Written without torch.distributions
    (argument validation there syncs with the host and cannot be captured in a CUDA graph)."""

    def __init__(self, in_features, n_actions, hidden_sizes, temperature=1.0):
        super().__init__(in_features, n_actions, hidden_sizes, activation_class=nn.ReLU)
        self.temperature = float(temperature)

    def forward(self, state):
        logits = super().forward(state)
        u = torch.rand_like(logits).clamp_(1e-20, 1.0 - 1e-7)
        relaxed = torch.softmax((logits - torch.log(-torch.log(u))) / self.temperature, dim=-1)
        hard = torch.nn.functional.one_hot(relaxed.argmax(dim=-1), logits.shape[-1]).to(logits.dtype)
        action = (hard - relaxed).detach() + relaxed
        log_prob = (action * torch.log_softmax(logits, dim=-1)).sum(-1, keepdim=True)
        return action, log_prob, logits
