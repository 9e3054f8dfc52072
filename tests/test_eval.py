"""Evaluation row of the testing scripts (EG:688-807 / EU:601-704): host formulas vs the oracle's literal NumPy
restatement (CPU), and the GPU confusion-count kernel vs NumPy (bit-exact integers)."""
import numpy as np
import pytest

from oracle import depgan_oracle as O


def _conf(fake, real):
    c = np.zeros((4, 4), np.int64)
    np.add.at(c, (real.reshape(-1).astype(np.int64), fake.reshape(-1).astype(np.int64)), 1)
    return c


@pytest.mark.parametrize("seed,probs", [(0, (0.9, 0.03, 0.04, 0.03)), (1, (0.25, 0.25, 0.25, 0.25)),
                                        (2, (1.0, 0.0, 0.0, 0.0)), (3, (0.5, 0.5, 0.0, 0.0)), (4, (0.7, 0.0, 0.0, 0.3))])
def test_evaluation_row_matches_oracle(seed, probs):
    from depgan_b200.postproc import evaluation_row
    rng = np.random.default_rng(seed)
    fake = rng.choice(4, size=(6, 32, 32), p=probs).astype(np.uint8)
    real = rng.choice(4, size=(6, 32, 32), p=probs[::-1] if seed % 2 else probs).astype(np.float64)  # brain_code is float
    for v1, v2, vo in [(10.0, 12.5, 11.0), (10.0, 10.0, 9.5), (10.0, 8.0, 9.0), (10.0, 8.0, 10.0), (0.0, 0.0, 0.0)]:
        got = evaluation_row(_conf(fake, real), v1, v2, vo)
        want = O.evaluation_row(fake, real, v1, v2, vo)
        assert len(got) == len(want) == 18
        assert all(float(a) == float(b) for a, b in zip(got, want)), (got, want)


def test_evaluation_row_known_answer():
    from depgan_b200.postproc import dice_scores_from_confusion, evaluation_row
    #            real: 0 0 1 1 2 3 3 3      fake: 0 1 1 2 2 3 3 0
    real = np.array([0, 0, 1, 1, 2, 3, 3, 3])
    fake = np.array([0, 1, 1, 2, 2, 3, 3, 0])
    d1, d2, d3, d4, d5, d6 = dice_scores_from_confusion(_conf(fake, real))
    s = 1e-7
    assert d1 == (2 * 1 + s) / (s + 2 + 2)      # shrink: 1 hit, 2 real, 2 fake
    assert d2 == (2 * 1 + s) / (s + 1 + 2)      # grow
    assert d3 == d6 == (2 * 2 + s) / (s + 3 + 2)  # stay
    assert d4 == (2 * 5 + s) / (s + 6 + 6)      # any WMH: real>0 at 6, fake>0 at 6, both at 5
    assert d5 == (2 * 3 + s) / (s + 3 + 4)      # changing: real {2,3,4}, fake {1,2,3,4}
    row = evaluation_row(_conf(fake, real), 5.0, 4.0, 4.5)   # shrinking, predicted shrinking
    assert row[:5] == [1, 0, 0, 1, 1] and row[8] == 0.25 and row[9] == 0.5
    row = evaluation_row(_conf(fake, real), 5.0, 5.0, 4.5)   # equal volumes count as progression (>=), missed
    assert row[:5] == [0, 1, 0, 0, 0]
    # empty classes: 0/0 -> smooth/smooth = 1
    z = np.zeros(8, np.int64)
    assert dice_scores_from_confusion(_conf(z, z)) == (1.0, 1.0, 1.0, 1.0, 1.0, 1.0)


@pytest.mark.gpu
@pytest.mark.parametrize("n", [0, 1, 1000003, 48 * 256 * 256])
def test_label_confusion_bit_exact(n):
    import torch
    from depgan_b200.postproc import label_confusion_device
    rng = np.random.default_rng(n % 97)
    fake = rng.integers(0, 4, size=n, dtype=np.uint8)
    real = rng.integers(0, 4, size=n, dtype=np.uint8)
    if n > 10:
        fake[3] = 7  # labels outside 0..3 are counted nowhere
        real[5] = 200
    got = label_confusion_device(torch.from_numpy(fake).cuda(), torch.from_numpy(real).cuda()).cpu().numpy()
    ok = (fake < 4) & (real < 4)
    want = _conf(fake[ok], real[ok])
    assert (got == want).all()


@pytest.mark.gpu
def test_evaluate_labels_end_to_end():
    from depgan_b200.postproc import evaluate_labels
    rng = np.random.default_rng(5)
    fake = rng.choice(4, size=(12, 64, 64), p=(0.9, 0.03, 0.04, 0.03)).astype(np.uint8)
    real = rng.choice(4, size=(12, 64, 64), p=(0.9, 0.04, 0.03, 0.03)).astype(np.float64)
    got = evaluate_labels(fake, real, 7.0, 8.0, 7.5)
    want = O.evaluation_row(fake, real, 7.0, 8.0, 7.5)
    assert all(float(a) == float(b) for a, b in zip(got, want))
