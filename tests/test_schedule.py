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
