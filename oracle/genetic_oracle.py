"""Restatement of the reference's mating rules with the random draws made explicit (TEST INFRASTRUCTURE ONLY).

EvolutionaryRacer/Mating.hpp:52-166 draws from unseeded std::random_device generators and its Network is an Eigen type
(Eigen is not installed here, so the file cannot be compiled): the pin is this line-by-line restatement, driven by an
INJECTED stream of draws.  Every function takes the uniforms / parent indices the reference would have drawn, in the
order it draws them, and follows its loops one coefficient at a time.

* mate2_selective            Mating.hpp:52-104   two uniform draws per coefficient, always: r1 decides "mutate" (< 0.1);
                                                 r2 is the mutation value ((r2 - 0.5) * 2) or picks the parent (< 0.75: the
                                                 superior one).  Superior = agent_1 only if its score is STRICTLY greater.
* choose_and_mate            Mating.hpp:113-166  sort by score (descending), keep 5; slot 0 = the best one's clone, slot 1 =
                                                 the best mated with itself; then (first, second != first) parents by
                                                 roulette until the colony is full.
Coefficients are visited in the flat order of the arrays handed in (the reference walks Eigen's storage order; the draws
are independent per coefficient, so the order only names which draw belongs to which coefficient)."""
import numpy as np

P_MUTATE = 0.1       # Mating.hpp:54
P_DOMINANT = 0.75    # Mating.hpp:55
K_PARENTS = 5        # Mating.hpp:118


def mate2_selective(w_a1, score_a1, w_a2, score_a2, draws):
    """one weight array of the offspring of (agent_1, agent_2); draws: f32[w.size, 2] = (r1, r2) per coefficient"""
    n1, n2 = (w_a1, w_a2) if score_a1 > score_a2 else (w_a2, w_a1)  # Mating.hpp:58-61: a tie makes agent_2 superior
    out = np.empty_like(n1)
    f1, f2, fo, d = n1.reshape(-1), n2.reshape(-1), out.reshape(-1), np.asarray(draws, dtype=np.float32).reshape(-1, 2)
    for i in range(f1.size):
        if d[i, 0] < P_MUTATE:
            fo[i] = np.float32((np.float64(d[i, 1]) - 0.5) * 2.0)  # scale from [0,1] to [-1,1]
        elif d[i, 1] < P_DOMINANT:
            fo[i] = f1[i]
        else:
            fo[i] = f2[i]
    return out


def choose_and_mate(w1, w2, scores, parents, draws_w1, draws_w2):
    """w1 f32[n, ...], w2 f32[n, ...], scores f32[n].  parents: the stream of roulette draws (indices into the sorted top
    five) in the order the reference consumes them.  draws_w1[k] / draws_w2[k]: the coefficient draws of the k-th call of
    mate2AgentsSelective (k = 0 is the best agent's self-mutation).  Returns (new_w1, new_w2, order[:5])."""
    n = len(scores)
    order = sorted(range(n), key=lambda i: -float(scores[i]))[:K_PARENTS]  # std::sort descending (ties: any order)
    top = [float(scores[i]) for i in order]
    new1, new2 = [w1[order[0]].copy()], [w2[order[0]].copy()]              # the clone, Mating.hpp:129
    new1.append(mate2_selective(w1[order[0]], top[0], w1[order[0]], top[0], draws_w1[0]))  # self-mutation, :130
    new2.append(mate2_selective(w2[order[0]], top[0], w2[order[0]], top[0], draws_w2[0]))
    it, k = iter(parents), 1
    while len(new1) < n:
        first = int(next(it))
        second = -1
        while second == -1 or second == first:  # "Prevent self-mutation", Mating.hpp:144-149
            second = int(next(it))
        new1.append(mate2_selective(w1[order[first]], top[first], w1[order[second]], top[second], draws_w1[k]))
        new2.append(mate2_selective(w2[order[first]], top[first], w2[order[second]], top[second], draws_w2[k]))
        k += 1
    return np.stack(new1), np.stack(new2), order
