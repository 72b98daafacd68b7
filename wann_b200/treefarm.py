"""TreeFarm: BASELINE config 5 -- many independent games per GPU, each contributing one leaf per
step, entirely device-resident: leaf boards -> wann_eval_expand (ranked priors + child boards +
win flags) -> pick a child per game -> the children ARE the next leaf batch.  No host round trip
inside a step; torch is used only for the per-game child selection (gather / multinomial).

This is a steady-state load generator and a demonstration of the child-construction path
(SURVEY 8f-1), not a re-implementation of WaNN's MCTS (out of scope)."""
import numpy as np

from . import synthetic


class TreeFarm(object):
    def __init__(self, evaluator, n_trees, device=None, seed=0, greedy=False):
        import torch
        self.torch = torch
        self.ev = evaluator
        self.n = int(n_trees)
        self.dev = torch.device("cuda", evaluator.device) if device is None else device
        self.greedy = greedy
        self.gen = torch.Generator(device=self.dev)
        self.gen.manual_seed(seed)
        self.seed, self.step_no = int(seed), 0
        self.opening = torch.from_numpy(synthetic.initial_squares()).to(self.dev)
        self.sq = self.opening[None].repeat(self.n, 1).contiguous()
        self.black = torch.zeros(self.n, dtype=torch.uint8, device=self.dev)
        self.rows = torch.arange(self.n, device=self.dev)
        self.games_finished = 0
        self.plies = 0

    def step(self):
        """One leaf per game: evaluate, expand, move.  Returns (idx, prior, n_legal, chosen slot)."""
        torch = self.torch
        idx, pri, cnt, kids, flg = self.ev.eval_expand(self.sq, self.black)
        if self.greedy:
            k = torch.zeros(self.n, dtype=torch.long, device=self.dev)       # policy move = slot 0
        else:
            w = pri.clamp_min(0) + (torch.arange(48, device=self.dev)[None] < cnt[:, None]) * 1e-12
            k = torch.multinomial(w, 1, generator=self.gen).squeeze(1)       # sample a legal child by prior
        child = kids[self.rows, k]
        won = (flg[self.rows, k] & 1).bool()
        stuck = cnt == 0                                                     # no legal move (cannot happen in play)
        reset = won | stuck
        self.sq = torch.where(reset[:, None], self.opening[None], child).contiguous()
        self.black = torch.where(reset, torch.zeros_like(self.black), self.black ^ 1)
        self.last_reset = reset
        self.plies += self.n
        return idx, pri, cnt, k

    def step_fused(self, want_outputs=False):
        """Same step as ONE C-ABI call (wann_tree_step): evaluation, child choice (counter-based RNG,
        reproducible on the host), move application and restart all stay on the device, boards are
        updated in place.  Returns (chosen idx, finished) device tensors when asked."""
        out = self.ev.tree_step(self.sq, self.black, greedy=self.greedy, seed=self.seed, step=self.step_no,
                                want_outputs=want_outputs)
        self.step_no += 1
        self.plies += self.n
        return out

    def run(self, steps):
        for _ in range(steps):
            self.step()
            self.games_finished += int(self.last_reset.sum())      # (syncs; use step() in timed loops)
        return self.plies
