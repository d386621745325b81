"""CPU: N2's mating rules pinned.  openkitchen_b200.genetic (batched torch ops) against oracle/genetic_oracle.py, the
coefficient-by-coefficient restatement of EvolutionaryRacer/Mating.hpp:52-166, on the SAME injected random draws:
identical colonies, bit for bit -- clone in slot 0, self-mutation in slot 1, roulette parents with second != first,
"superior" = strictly greater score (a tie favours agent_2), 10 % mutation to (r2 - 0.5) * 2, 75 % superior parent."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from openkitchen_b200.genetic import K_PARENTS, GeneticPopulation  # noqa: E402
from oracle.genetic_oracle import choose_and_mate, mate2_selective  # noqa: E402


def _bits(a, b):
    return np.array_equal(np.asarray(a, np.float32).view(np.uint32), np.asarray(b, np.float32).view(np.uint32))


@pytest.mark.parametrize("tie", [False, True])
def test_colony_matches_the_reference_rules_on_injected_draws(tie):
    n, rays, hidden = 23, 15, 30  # Network.hpp:92-95 sizes: (15 + 2) x 30 and 30 x 6
    rng = np.random.default_rng(3 + tie)
    pop = GeneticPopulation(n, rays, device="cpu", generator=torch.Generator().manual_seed(0), hidden=hidden)
    w1, w2 = pop.w1.numpy().copy(), pop.w2.numpy().copy()
    scores = rng.permutation(n).astype(np.float32) * 10.0
    if tie:  # two of the parents share a score: "superior" then is agent_2, whichever order the roulette names them in
        top = np.argsort(-scores)[:K_PARENTS]
        scores[top[2]] = scores[top[1]]
    m = n - 2
    first = rng.integers(0, K_PARENTS, m)
    second = (first + rng.integers(1, K_PARENTS, m)) % K_PARENTS  # != first
    stream = [v for pair in zip(first, second) for v in pair]
    # with rejected draws in the stream: the reference loops until second != first
    stream_rej = []
    for f, s in zip(first, second):
        stream_rej += [f] + ([f] if rng.random() < 0.3 else []) + [s]
    d1 = rng.random((m + 1, rays + 2, hidden, 2)).astype(np.float32)
    d2 = rng.random((m + 1, hidden, 6, 2)).astype(np.float32)
    want1, want2, order = choose_and_mate(w1, w2, scores, stream_rej, d1, d2)
    assert _bits(want1, choose_and_mate(w1, w2, scores, stream, d1, d2)[0])
    # the product, on the same draws
    sc = torch.from_numpy(scores)
    top_val, top_idx = torch.sort(sc, descending=True, stable=True)
    top_val, top_idx = top_val[:K_PARENTS], top_idx[:K_PARENTS]
    if not tie:
        assert top_idx.tolist() == order
    pw1, pw2 = pop.w1[top_idx], pop.w2[top_idx]
    # with tied scores std::sort may order the tied parents either way: use the oracle's order for the comparison
    if tie:
        top_idx = torch.tensor(order)
        top_val, pw1, pw2 = sc[top_idx], pop.w1[top_idx], pop.w2[top_idx]
    got1, got2 = pop.mate_from_draws(pw1, pw2, top_val, torch.from_numpy(first), torch.from_numpy(second),
                                     torch.from_numpy(d1), torch.from_numpy(d2))
    assert got1.shape == (n, rays + 2, hidden) and got2.shape == (n, hidden, 6)
    assert _bits(got1.numpy(), want1) and _bits(got2.numpy(), want2)
    assert _bits(got1[0].numpy(), w1[order[0]])  # the clone


def test_tie_makes_the_second_agent_superior():
    a, b = np.full((2, 3), 1.0, np.float32), np.full((2, 3), 2.0, np.float32)
    draws = np.stack([np.full(6, 0.5, np.float32), np.full(6, 0.2, np.float32)], axis=1)  # no mutation, take the superior one
    assert (mate2_selective(a, 5.0, b, 5.0, draws) == 2.0).all()  # tie: agent_2
    assert (mate2_selective(a, 6.0, b, 5.0, draws) == 1.0).all()
    pop = GeneticPopulation(4, 1, device="cpu", generator=torch.Generator().manual_seed(0), hidden=2)
    pw1 = torch.stack([torch.full((3, 2), float(k)) for k in range(5)])
    pw2 = torch.stack([torch.full((2, 6), float(k)) for k in range(5)])
    vals = torch.tensor([9.0, 7.0, 7.0, 3.0, 1.0])
    d1 = torch.tensor([0.5, 0.2]).expand(1, 3, 2, 2).contiguous()
    d2 = torch.tensor([0.5, 0.2]).expand(1, 2, 6, 2).contiguous()
    n1, _ = pop.mate_from_draws(pw1, pw2, vals, torch.tensor([1]), torch.tensor([2]), d1, d2, with_elite=False)
    assert (n1 == 2.0).all()  # scores tie: the SECOND named parent is superior
    n1, _ = pop.mate_from_draws(pw1, pw2, vals, torch.tensor([2]), torch.tensor([1]), d1, d2, with_elite=False)
    assert (n1 == 1.0).all()


def test_mate_statistics_and_shapes():
    n, rays = 400, 15
    pop = GeneticPopulation(n, rays, device="cpu", generator=torch.Generator().manual_seed(1))
    best = pop.w1[17].clone()
    scores = torch.arange(n, dtype=torch.float32)
    scores[17] = 1e6
    top_val, top_idx = pop.mate(scores)
    assert top_idx[0] == 17 and torch.equal(pop.w1[0], best)
    changed = (pop.w1[1] != best).float().mean()
    assert 0.05 < changed < 0.16  # ~10 % of the self-mutation's coefficients are re-drawn
    assert pop.w1.shape == (n, rays + 2, 30) and pop.w2.shape == (n, 30, 6)
    assert pop.w1.abs().max() <= 1.0
