"""TEST INFRASTRUCTURE (oracle): the reference's search tree restated over flat arrays.

A CPU restatement, node statistics in parallel Python lists instead of dict nodes, of what ONE thread
of the reference's search does ("array-based tree replacing dict nodes"):

  run_simulation            expansion_MCTS_functions.pyx:300-350 (run_MCTS_with_expansions_simulation,
                            play_MCTS_game_with_expansions, expand_leaf_node, select_UCT_child)
  choose / find_best_UCT    tree_search_utils.pyx:251-403, get_UCT :405-467
  expand_descendants        tree_builder.pyx:122-215, do_updates_in_the_same_process :294-340,
                            update_parent :344-355, enumerate_update_and_prune :428-445,
                            get_child :530-536, assign_children :538-620
  seeds                     oracle.seed_children (update_child, check_for_winning_move, threat branch,
                            evaluation_function) -- already pinned by ref_seed_golden.npz
  statistics / solver       update_tree_wins/_losses :43-66, update_win_statuses & co :70-132,
                            backpropagate_num_checked_children :134-158, thread counters :161-242

Pinned by tests/golden/ref_tree_golden.npz: whole trees grown by the compiled reference,
single-threaded, checkpointed after 1..40 simulations, reproduced node for node.
The network is any callable (squares int8[64], black) -> float32[155]."""
import numpy as np

from . import oracle as O


class ArrayTree(object):
    F = ("parent", "index", "black", "visits", "wins", "gv", "gw", "ws", "gameover", "checked", "being_checked",
         "threads", "first", "count", "height", "prior", "uct")

    def __init__(self, squares, black_to_move, policy, height=0):
        for f in self.F:
            setattr(self, f, [])
        self.board = []
        self.policy = policy
        self.evaluations = 0
        self.root = self._new_node(-1, 0, bool(black_to_move), np.asarray(squares, np.int8).copy(), height, 0.0, 1.0)

    # ------------------------------------------------------------------ node storage (TreeNode.pyx:7-50)
    def _new_node(self, parent, index, black, board, height, prior, uct):
        for f, v in zip(self.F, (parent, index, black, 0, 0, 0, 0, -1, False, False, False, 0, -1, -1, height, prior, uct)):
            getattr(self, f).append(v)
        self.board.append(board)
        return len(self.parent) - 1

    def children(self, n):
        return range(self.first[n], self.first[n] + self.count[n]) if self.count[n] >= 0 else None

    # ------------------------------------------------------------------ statistics (tree_search_utils.pyx:43-66)
    def update_tree_wins(self, n, amount=1, gameover=False):
        win = True                                      # node win => parent loss => grandparent win ...
        while n >= 0:
            if win:
                self.wins[n] += amount
            self.visits[n] += amount
            if gameover:
                if win:
                    self.gw[n] += amount
                self.gv[n] += amount
            n, win = self.parent[n], not win

    def update_tree_losses(self, n, amount=1, gameover=False):
        self.visits[n] += amount
        if gameover:
            self.gv[n] += amount
        if self.parent[n] >= 0:
            self.update_tree_wins(self.parent[n], amount, gameover)

    # ------------------------------------------------------------------ solver (tree_search_utils.pyx:70-132)
    def update_win_statuses(self, n, status):
        if status == 0:
            self.update_tree_losses(n, 1, gameover=True)
        self.ws[n] = status
        if self.parent[n] >= 0:
            self.update_win_status_from_children(self.parent[n])

    def update_win_status_from_children(self, n):
        threshold = float(np.float32(1.001))
        kids = self.children(n)
        all_s, cons = [], []
        if kids is not None:
            for c in kids:
                if self.uct[c] > threshold:
                    cons.append(self.ws[c])
                all_s.append(self.ws[c])
        if 0 not in all_s and cons:
            all_s = cons
        if 0 in all_s:
            self.update_win_statuses(n, 1)
        elif 1 in all_s and -1 not in all_s:
            self.update_win_statuses(n, 0)

    def backpropagate_num_checked_children(self, n):         # :134-158
        while n >= 0:
            kids = self.children(n)
            if sum(1 for c in kids if self.checked[c]) == len(kids):
                self.checked[n] = True
                n = self.parent[n]
            else:
                self.checked[n] = False
                n = -1

    # ------------------------------------------------------------------ thread counters (:161-242)
    def _crement_threads(self, n, amount):
        if amount == 0:
            self.threads[n] = 0
        else:
            self.threads[n] += amount
        if self.parent[n] >= 0:
            self.update_num_children_being_checked(self.parent[n])

    def update_num_children_being_checked(self, n):
        kids = self.children(n)
        if kids is None:
            return
        cnt = sum(1 for c in kids if (self.threads[c] > 0 and self.count[c] < 0) or self.being_checked[c] or self.ws[c] != -1)
        previous = self.being_checked[n]
        self.being_checked[n] = cnt == len(kids)
        if previous != self.being_checked[n] and self.parent[n] >= 0:
            self.update_num_children_being_checked(self.parent[n])

    # ------------------------------------------------------------------ selection (:251-467)
    def get_uct(self, c, parent_visits):
        norm_wins, norm_visits = int(self.wins[c]), int(self.visits[c])
        true_wins, total_go = int(self.gw[c]), int(self.gv[c])
        nn_prob = max(0.10, float(self.uct[c]) - 1.0) if not isinstance(self.uct[c], np.float32) else \
            max(0.10, float(np.float32(self.uct[c] - np.float32(1))))
        if total_go > 10 and self.black[c] != self.black[self.root]:
            rate = (total_go - true_wins) / total_go + 1
        else:
            rate = (norm_visits - norm_wins) / norm_visits
        sibling_visits = max(1, parent_visits - norm_visits)
        return rate + 1.0 * nn_prob * (np.sqrt(float(sibling_visits)) / (1 + norm_visits))

    def find_best_uct_child(self, n):
        kids = list(self.children(n))
        ok_threads = lambda c: self.threads[c] <= 0 or self.count[c] >= 0        # noqa: E731
        open_ = lambda c: self.ws[c] == -1 and not self.being_checked[c]           # noqa: E731
        if self.ws[n] != -1:
            viable = [c for c in kids if open_(c) and ok_threads(c)]
            if not viable:
                viable = [c for c in kids if not (self.checked[c] or self.being_checked[c]) and ok_threads(c)]
                if not viable:
                    viable = [c for c in kids if not self.checked[c] and ok_threads(c)]
        else:
            if self.black[n] == self.black[self.root]:
                best_prob = float(self.uct[kids[0]]) - 1 if self.height[n] < 60 else .01
                if best_prob < .1:
                    first = second = 1.04
                elif best_prob < .2:
                    first, second = 1.10, 1.05
                elif best_prob < .3:
                    first, second = 1.15, 1.10
                else:
                    first, second = 1.20, 1.15
            else:
                first, second = 1.1, 1.01
            first, second = float(np.float32(first)), float(np.float32(second))      # `float first_threshold` (C float)
            viable = [c for c in kids if open_(c) and self.uct[c] > first and ok_threads(c)]
            if not viable:
                if n == self.root:
                    viable = [c for c in kids if open_(c) and self.uct[c] > second and
                              (self.gw[c] <= 1000 or self.gw[c] * 2 < self.gv[c]) and ok_threads(c)]
                    if not viable:
                        viable = [c for c in kids if open_(c) and self.uct[c] > second and ok_threads(c)]
                        if not viable:
                            viable = [c for c in kids if open_(c) and ok_threads(c)]
                else:
                    viable = [c for c in kids if open_(c) and ok_threads(c)]
        if not viable:
            return -1
        pv = self.visits[n]
        best, best_v = -1, None
        for c in viable:                                 # max(): the FIRST maximum wins
            v = self.get_uct(c, pv)
            if best_v is None or v > best_v:
                best, best_v = c, v
        return best

    def choose_uct_or_best_child(self, n):
        if self.count[n] == 1:
            return self.first[n]
        return self.find_best_uct_child(n)

    # ------------------------------------------------------------------ expansion (tree_builder.pyx)
    def _expand_one(self, n):
        """update_parent -> enumerate_update_and_prune -> get_child x k -> assign_children.
        Returns the list the reference's assign_children returns (next-frontier candidates)."""
        sq, blk = self.board[n], self.black[n]
        probs = np.asarray(self.policy(sq, blk), np.float32)
        self.evaluations += 1
        mask = O.legal_mask(sq[None], np.array([blk], np.uint8))
        idx, pri, cnt = O.ranked_children(probs[None], mask)
        k = int(cnt[0])
        stats, flags, _ = O.seed_children(sq[None], np.array([blk], np.uint8), idx, pri, cnt)
        first = len(self.parent)
        kids = []
        # get_child: every child is created (and a winning child's loss propagated) before the sort
        for j in range(k):
            f = int(flags[0, j])
            p32 = np.float32(pri[0, j])
            uct = np.float32(np.float32(1) + p32) if f & O.SEED_CLOSE_TO_HOME else 1.0 + float(p32)
            c = self._new_node(n, int(idx[0, j]), not blk, O.apply_pov_move(sq, int(blk), int(idx[0, j])),
                               self.height[n] + 1, p32, uct)
            kids.append(c)
            if f & O.SEED_WIN:                           # set_game_over_values + update_tree_losses(child)
                self.gameover[c] = self.checked[c] = True
                self.ws[c] = 0
                self.gv[c] = self.visits[c] = 65536
                self.update_tree_losses(c, 1, gameover=True)
        # update_child seeds (everything but the propagation, which is done explicitly here)
        threat = any(int(flags[0, j]) & (O.SEED_GAME_SAVING | O.SEED_DOOMED) for j in range(k)) or \
            bool((O.pov_squares(sq[None], np.array([blk], np.uint8))[0][8:16] == 2).any())
        # assign_children: decrement, sort (already ranked), threat branch or eval_children
        self._crement_threads(n, -1)
        returned = []
        if k > 0 and self.count[n] < 0:
            ev_l = ev_w = 0
            for j, c in enumerate(kids):
                f = int(flags[0, j])
                v, w, gv, gw = (int(x) for x in stats[0, j])
                if threat:
                    if f & O.SEED_GAME_SAVING:
                        self.visits[c], self.wins[c], self.gv[c], self.gw[c] = v, w, gv, gw
                        returned.append(c)
                    elif f & O.SEED_DOOMED:              # set_game_over_values_1_step_lookahead + update_tree_wins(child)
                        self.visits[c] = self.wins[c] = self.gw[c] = self.gv[c] = 65536
                        self.ws[c] = 1
                        self.update_tree_wins(c, 1, gameover=True)
                    # a winning child keeps its game-over values
                elif not f & O.SEED_WIN:
                    self.visits[c], self.wins[c], self.gv[c], self.gw[c] = v, w, gv, gw
                    if f & O.SEED_EVAL_ONE:
                        ev_l += 1
                    else:
                        ev_w += 1
                else:
                    if f & O.SEED_EVAL_ONE:
                        ev_l += 1
                    else:
                        ev_w += 1
            self.first[n], self.count[n] = first, k
            if not threat:
                returned = list(kids)
                self.update_tree_losses(n, ev_l)         # eval_children
                self.update_tree_wins(n, ev_w)
            self.update_win_status_from_children(n)
            self.backpropagate_num_checked_children(n)
            self.update_num_children_being_checked(n)
        return returned

    def expand_descendants(self, leaf, depth, depth_limit, no_root=False):
        """tree_builder.pyx:122-215 for a one-node frontier (what the tree always passes).
        no_root: sim_info['root'] is None (init_new_root runs before the root is assigned): the frontier
        filter then ignores win statuses (:142-143,156-159)."""
        frontier = []
        while depth < depth_limit:
            if not no_root and self.ws[self.root] == -1:
                if self.count[leaf] < 0 and self.ws[leaf] == -1:
                    frontier.append(leaf)
            elif self.count[leaf] < 0 and not self.gameover[leaf]:
                frontier.append(leaf)
            if 0 < len(frontier) < 256:
                depth_limit = 800
                nxt = []
                for p in frontier:                       # do_updates_in_the_same_process
                    if self.threads[p] <= 1 and self.count[p] < 0:
                        if self.threads[p] <= 0:
                            self._crement_threads(p, 1)
                        if self.count[p] < 0:            # update_parent
                            kids = self._expand_one(p)
                            if depth + 1 < depth_limit and kids:
                                c0 = kids[0]
                                if self.ws[c0] == -1 and self.threads[c0] <= 0 and not self.being_checked[c0]:
                                    nxt.append(c0)
                        else:
                            self._crement_threads(p, -1)
                    else:
                        self._crement_threads(p, -1)
                frontier = nxt
                depth += 1
                if depth < depth_limit:
                    for c in frontier:
                        self._crement_threads(c, 1)
            else:
                depth = depth_limit
                self._crement_threads(leaf, 0)

    # ------------------------------------------------------------------ one simulation
    def run_simulation(self):
        root = self.root
        if self.checked[root]:
            return True
        self._play(root)
        done = self.checked[root] or self.ws[root] != -1
        if done and self.count[root] < 0:
            self.expand_descendants(root, 0, 1)
        return done

    def _play(self, n):
        while True:
            if self.gameover[n]:
                return
            if self.count[n] < 0:
                if self.threads[n] <= 0:
                    self._crement_threads(n, 1)
                    self.expand_descendants(n, 0, 5)
                return
            while n >= 0 and self.count[n] >= 0:          # select_UCT_child
                self.threads[n] = 0
                n = self.choose_uct_or_best_child(n)
            if n < 0:
                return

    # ------------------------------------------------------------------ root reuse between moves
    def advance_root(self, move_index):
        """assign_root's "Reused old tree" branch (expansion_MCTS_functions.pyx:236-247,253-257) followed by
        reset_thread_flag (:197-207): the child reached by `move_index` becomes the root of the next
        search and the thread bookkeeping of its subtree is cleared.  (The reference also cuts the tree
        above the new root's grandparent, :259-266 -- memory housekeeping with no effect below the root.)"""
        if self.count[self.root] < 0:
            # init_new_root (tree_builder.pyx:809-897): the previous node was never expanded -- expand it
            # now (threads_checking_node = 1, expand_descendants([parent], 0, 1) with sim_info['root'] still
            # None), then continue as below.  (Its closing backpropagate_num_checked_children /
            # update_win_status_from_children(new_subtree=True) only touch the parent and above.)
            self.threads[self.root] = 1
            self.expand_descendants(self.root, 0, 1, no_root=True)
        kid = [c for c in self.children(self.root) if self.index[c] == move_index]
        if not kid:
            raise KeyError("the root has no child for move index %d" % move_index)
        self.root = kid[0]
        stack = [self.root]
        while stack:
            n = stack.pop()
            self.threads[n] = 0
            self.being_checked[n] = False
            if self.count[n] >= 0:
                stack.extend(self.children(n))

    # ------------------------------------------------------------------ the move the search plays
    def best_move(self):
        """randomly_choose_a_winning_move (tree_search_utils.pyx:469-486; deterministic despite its name:
        it returns best_nodes[0]) with check_for_forced_win :559-582 and choose_best_true_loser :493-557.
        -> node of the chosen child (its move index is self.index[node]).  The expressions are kept as the
        reference writes them: a capture close to home carries a numpy float32 UCT_multiplier, and numpy's
        scalar promotion then decides in which precision each sum and comparison happens."""
        kids = list(self.children(self.root))
        over = [c for c in kids if self.gameover[c]]
        if over:
            return over[0]
        lost = [c for c in kids if self.ws[c] == 0]
        if lost:
            return lost[0]
        return self._choose_best_true_loser(kids)

    def _choose_best_true_loser(self, kids):
        prob_scaling = 3
        best_child = kids[0]
        action_value_threshold = self.gw[best_child] / 2
        top = self.uct[best_child]
        if top < 1.1:
            first = second = 1.04
        elif top < 1.2:
            first, second = 1.10, 1.05
        elif top < 1.3:
            first, second = 1.15, 1.10
        else:
            first, second = 1.20, 1.15
        losses = lambda c: self.gv[c] - self.gw[c]          # noqa: E731
        filtered = [c for c in kids if (self.uct[c] > first or losses(c) > action_value_threshold) and self.ws[c] != 1]
        if not filtered:
            filtered = [c for c in kids if (self.uct[c] > second or losses(c) > action_value_threshold) and self.ws[c] != 1]
        if not filtered:
            filtered = [c for c in kids if self.ws[c] != 1]
        if not filtered:
            return kids[0]
        best, best_losses = None, None
        for c in filtered:                                   # max(): the first maximum
            if best is None or losses(c) > best_losses:
                best, best_losses = c, losses(c)
        highest = best_losses
        best_rate = best_losses / max(1, self.gv[best]) + (self.uct[best] - 1) / prob_scaling
        for c in filtered:
            true_losses = losses(c)
            loss_rate = (true_losses / max(1, self.gv[c])) + (self.uct[c] - 1) / prob_scaling
            if loss_rate > best_rate and loss_rate - best_rate > 0.05 and highest != 0 and true_losses / highest > .30 \
                    and self.ws[c] != 1:
                best, best_rate = c, loss_rate
        return best

    # ------------------------------------------------------------------ comparison with the golden dump
    def dump(self):
        """Same BFS layout as tests/golden/make_tree_golden.py: children contiguous, in tree order."""
        order, rows, i = [(self.root, -1)], [], 0
        while i < len(order):
            n, par = order[i]
            rows.append((par, self.index[n], int(self.black[n]), self.visits[n], int(self.wins[n]), self.gv[n],
                         int(self.gw[n]), self.ws[n], int(self.gameover[n]), int(self.checked[n]),
                         int(self.being_checked[n]), self.threads[n], self.count[n], self.height[n]))
            if self.count[n] >= 0:
                order.extend((c, i) for c in self.children(n))
            i += 1
        return np.array(rows, np.int64)
