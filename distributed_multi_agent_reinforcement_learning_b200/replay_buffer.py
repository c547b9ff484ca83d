"""Episode buffers — mirror of the reference `DHGN/replay_buffer.py` (ReplayBuffer, BigBuffer): same keys, shapes and
methods, tensors resident on the GPU.  The batched engine does not go through these (it trains straight from the
time-major `RolloutArena` / `TrainBatch`); they exist so that reference-style callers keep working and so that a
reference buffer dict can be fed to `MAPPO.train`."""
import torch

from . import maps


class ReplayBuffer:
    KEYS = ("p_state", "e_state", "o_state", "p_adj", "e_adj", "o_adj", "actor_historical_embedding",
            "critic_historical_embedding", "v_n", "a_n", "a_logprob_n", "r", "active")

    def __init__(self, cfg, device=None):
        self.episode_limit = cfg.env.max_steps
        self.batch_size = cfg.algo.sample_epi_num
        self.episode_num, self.max_episode_len = 0, 0
        self.device = torch.device(device if device is not None else cfg.algo.worker_device)
        self.max_p_num, self.max_e_num = cfg.env.num_defender, cfg.env.num_attacker
        self.max_o_num = cfg.map.num_max_obstacle
        self.p_dim = self.e_dim = self.o_dim = cfg.env.state_dim
        self.embedding_size, self.depth = cfg.algo.embedding_dim, cfg.algo.depth
        self.buffer = None

    def reset_buffer(self):
        B, T, P, Ev, O, D, E = (self.batch_size, self.episode_limit, self.max_p_num, self.max_e_num, self.max_o_num,
                                self.depth, self.embedding_size)
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=self.device)
        self.buffer = dict(
            p_state=z(B, T, P, self.p_dim), e_state=z(B, T, Ev, self.e_dim), o_state=z(B, T, O, self.o_dim),
            p_adj=z(B, T, P, P), e_adj=z(B, T, P, Ev), o_adj=z(B, T, P, O),
            actor_historical_embedding=z(B, T + D, P, E), critic_historical_embedding=z(B, T + D, P, E),
            v_n=z(B, T + 1, P), a_n=z(B, T, P), a_logprob_n=z(B, T, P), r=z(B, T, P), active=z(B, T, P))

    def store_transition(self, num_episode, episode_step, p_state, e_state, o_state, p_adj, e_adj, o_adj,
                         actor_historical_embedding, critic_historical_embedding, v_n, a_n, a_logprob_n, r, active):
        b, k, t = self.buffer, num_episode, episode_step
        pn, en, on = len(p_state), len(e_state), len(o_state)
        b["p_state"][k, t, :pn] = p_state
        b["e_state"][k, t, :en] = e_state
        b["o_state"][k, t, :on] = o_state
        b["p_adj"][k, t, :pn, :pn] = p_adj
        b["e_adj"][k, t, :pn, :en] = e_adj
        b["o_adj"][k, t, :pn, :on] = o_adj
        b["actor_historical_embedding"][k, t + self.depth, :pn] = actor_historical_embedding
        b["critic_historical_embedding"][k, t + self.depth, :pn] = critic_historical_embedding
        b["v_n"][k, t, :pn] = v_n
        b["a_n"][k, t, :pn] = a_n
        b["a_logprob_n"][k, t, :pn] = a_logprob_n
        b["r"][k, t, :pn] = r
        b["active"][k, t, :pn] = active

    def store_last_value(self, num_episode, episode_step, v_n):
        self.buffer["v_n"][num_episode, episode_step, :len(v_n)] = v_n

    def store_episode(self, num_episode, arena, tb, env_index=0):
        """Copies one env's episode out of a time-major arena / TrainBatch into slot `num_episode` (reference layout)."""
        b, k, e = self.buffer, num_episode, env_index
        T = tb.T
        b["p_state"][k, :T] = tb.p[:, e]
        b["e_state"][k, :T, 0] = tb.e[:, e]
        n_o = tb.oxy.shape[1]
        b["o_state"][k, :T, :n_o, :2] = tb.oxy[e].unsqueeze(0)
        N, O = tb.N, self.max_o_num
        dev = self.device
        unpack = lambda w, n: ((w.unsqueeze(-1) >> torch.arange(32, device=dev, dtype=torch.int32)) & 1).flatten(-2)[..., :n].float()
        b["p_adj"][k, :T] = unpack(tb.p_adj_bits[:, e], N)
        b["e_adj"][k, :T, :, 0] = tb.e_adj[:, e].float()
        b["o_adj"][k, :T] = unpack(tb.o_adj_bits[:, e], O)
        b["actor_historical_embedding"][k] = tb.hist_a[:, e]
        b["critic_historical_embedding"][k] = tb.hist_c[:, e]
        b["v_n"][k] = tb.v[:, e]
        b["a_n"][k, :T] = tb.a[:, e]
        b["a_logprob_n"][k, :T] = tb.logp[:, e]
        b["r"][k, :T] = tb.r[:, e]
        b["active"][k, :T] = tb.active[:, e]


class BigBuffer:
    def __init__(self):
        self.buffer = None

    def get_training_data(self, device):
        for key in self.buffer:
            self.buffer[key] = self.buffer[key].to(device)
        return self.buffer

    def reset(self):
        self.buffer = None

    def concat_buffer(self, mini_buffer):
        if self.buffer is None:
            self.buffer = dict(mini_buffer.buffer)
        else:
            for key in self.buffer:
                self.buffer[key] = torch.cat([self.buffer[key], mini_buffer.buffer[key]], dim=0)
