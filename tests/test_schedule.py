"""The epoch schedule of the DEP-GAN training loop (TG:790-894) against a literal transcription of the reference's
while-loops."""
import pytest

from depgan_b200.trainer import ScalarLog, epoch_schedule


def _reference_events(batches, gen_iterations, Diters):
    """TG:790-894 with the K.function calls replaced by event records."""
    ev = []
    i = 0
    ii = 0
    real_data = None
    while i < batches:
        if gen_iterations < 25 or gen_iterations % 500 == 0:
            _Diters = 100
            _Diters_dem = 100
        else:
            _Diters = Diters
            _Diters_dem = Diters
        j = 0
        jj = 0
        while j < _Diters and i < batches:
            j += 1
            real_data = i
            i += 1
            ev.append(("y2", real_data))
        while jj < _Diters_dem and ii < batches:
            jj += 1
            real_data = ii
            ii += 1
            ev.append(("dem", real_data))
        ev.append(("gen", real_data, gen_iterations))
        gen_iterations += 1
    return ev


@pytest.mark.parametrize("batches,g0,D", [(0, 0, 5), (3, 0, 5), (100, 0, 5), (250, 24, 5), (57, 30, 5), (1200, 498, 5),
                                          (40, 1000, 3), (11, 26, 5)])
def test_epoch_schedule_matches_reference_loops(batches, g0, D):
    assert list(epoch_schedule(batches, g0, D)) == _reference_events(batches, g0, D)


def test_scalar_log(tmp_path):
    lg = ScalarLog(tmp_path / "log.csv")
    lg.log_scalar("a", 1.5, 0)
    lg.log_scalar("a", 2.5, 1)
    assert lg.series["a"] == [(0, 1.5), (1, 2.5)]
    assert (tmp_path / "log.csv").read_text().splitlines() == ["a,0,1.5", "a,1,2.5"]


# ---- against the EXECUTED reference loop (tests/golden/reference_vectors.npz, made by executing TG:778-894 with
# recording stand-ins, see tests/golden/make_reference_vectors.py::training_loop_trace) --------------------------------
import json  # noqa: E402
from pathlib import Path  # noqa: E402

import numpy as np  # noqa: E402

_G = np.load(Path(__file__).parent / "golden" / "reference_vectors.npz", allow_pickle=False)
_SCENARIOS = ["warmup_to_steady", "every_500th", "from_scratch"]


@pytest.mark.parametrize("tag", _SCENARIOS)
def test_epoch_schedule_matches_the_executed_reference_loop(tag):
    d = json.loads(str(_G["loop/" + tag]))
    want = [tuple(e) for e in d["events"] if e[0] in ("y2", "dem", "gen")]
    got, g = [], d["g0"]
    for _ in range(d["niter"]):
        ev = list(epoch_schedule(d["batches"], g, 5))
        g += sum(1 for e in ev if e[0] == "gen")
        got += ev
    assert got == want and g == d["gen_iterations_end"]


class _Net:
    def __init__(self, ev):
        self.ev = ev

    def predict(self, x, **kw):
        return np.zeros((1,), np.float32)

    def save(self, path):
        self.ev.append(["save"])


class _Log:
    def __init__(self, ev):
        self.ev = ev

    def log_scalar(self, tag, value, step):
        self.ev.append(["log", tag, int(step)])

    def log_images(self, tag, images, step, *a):
        self.ev.append(["img", tag, int(step)])


@pytest.mark.parametrize("tag", _SCENARIOS)
def test_fit_reproduces_the_executed_reference_loop_event_for_event(tag):
    """DepGanTrainer.fit with its step functions replaced by recorders: the same mini-batch for every critic update and
    generator update, the same TensorBoard tags with the same step counters in the same order, validation every 10th and
    image summaries every 500th generator iteration, a save after every generator update."""
    from depgan_b200.trainer import DepGanTrainer
    d = json.loads(str(_G["loop/" + tag]))
    bs = d["batchSize"]
    ev = []
    tr = object.__new__(DepGanTrainer)   # no GPU: only the loop is exercised
    tr.gen_iterations = d["g0"]
    tr.Dy2, tr.G = _Net(ev), _Net(ev)

    def critic(kind):
        def f(inputs):
            r2, r1, noise, ep = inputs
            assert noise.shape == (bs, 4, 1) and ep.shape == (bs, 1, 1, 1)
            ev.append([kind, int(r1.ravel()[0]) // bs])
            return 0.25, 0.75
        return f

    def gen_iteration(cy, cd, r1, r2, noises):
        assert noises.shape == (10, bs, 4, 1) and noises.dtype == np.float32 and not cy and not cd
        ev.append(["gen", int(r1.ravel()[0]) // bs, tr.gen_iterations])
        tr.gen_iterations += 1
        return 0, [0.0] * 10, [1.0, 2.0, 3.0, 4.0, 5.0, 6.0]

    tr.netD_y2_train, tr.netD_dem_train, tr.gen_iteration = critic("y2"), critic("dem"), gen_iteration
    data = np.arange(d["batches"] * bs + 1, dtype=np.float32).reshape(-1, 1, 1, 1)
    val = (np.zeros((3, 2, 2, 1), np.float32), np.zeros((3, 2, 2, 1), np.float32))
    tr.fit(data, data.copy(), niter=d["niter"], batchSize=bs, Diters=5, noiseSize=4, k_noise=10, val=val,
           fixed_noise=np.zeros((3, 4, 1), np.float32), logger=_Log(ev), save_path="x.h5", save_every=1, shuffle=False)
    assert tr.gen_iterations == d["gen_iterations_end"]
    assert len(ev) == len(d["events"])
    for i, (a, b) in enumerate(zip(ev, d["events"])):
        assert a == b, (i, a, b)
