"""Host logic of the generator worker (core/models.py:_GeneratorFeed = Keras-1 GeneratorEnqueuer for one worker,
train.py:213-217): fetches exactly the batches the call consumes, keeps their order, surfaces generator errors,
and nb_worker=0 pulls in line."""
import itertools
import time

import numpy as np
import pytest

from asr_study_b200.core.models import _GeneratorFeed


def _gen(batch, log, fail_at=None):
    for k in itertools.count():
        if fail_at is not None and k == fail_at:
            raise RuntimeError("generator broke at batch %d" % k)
        log.append(k)
        x = np.full((batch, 5, 3), float(k), np.float32)
        yield ([x, None, np.full(batch, 5)], [np.zeros(batch), None])


@pytest.mark.parametrize("nb_worker", [0, 1])
def test_feed_fetches_exactly_what_is_consumed_and_keeps_order(nb_worker):
    log = []
    g = _gen(4, log)
    feed = _GeneratorFeed(g, total_samples=20, max_q_size=2, nb_worker=nb_worker, device="cpu")
    got = []
    for _ in range(5):
        x, y = feed.get()
        got.append(int(x[0][0, 0, 0]))
        time.sleep(0.01)
    feed.close()
    assert got == [0, 1, 2, 3, 4]
    assert log == [0, 1, 2, 3, 4]                       # never ahead of the 20 samples this call consumes
    assert next(g)[0][0][0, 0, 0] == 5.0                 # the generator's position is what a synchronous loop leaves


def test_feed_bounded_queue_does_not_run_ahead_of_max_q_size():
    log = []
    feed = _GeneratorFeed(_gen(2, log), total_samples=200, max_q_size=3, nb_worker=1, device="cpu")
    time.sleep(0.3)
    assert len(log) <= 3 + 1                            # queue of 3 + the batch the worker holds while the queue is full
    feed.get()
    feed.close()


def test_feed_surfaces_generator_errors():
    feed = _GeneratorFeed(_gen(2, [], fail_at=1), total_samples=100, max_q_size=4, nb_worker=1, device="cpu")
    feed.get()
    with pytest.raises(RuntimeError, match="generator broke"):
        feed.get()
    feed.close()


def test_fit_generator_loop_consumes_exactly_the_epochs_and_aggregates_on_read_back(monkeypatch):
    """The epoch loop of CTCModel.fit_generator with the device step replaced by a stub (CPU): it pulls exactly
    nb_epoch * samples_per_epoch utterances through the worker, weights the per-batch metrics by batch size, calls the
    callbacks once per epoch, and stops the worker when the step raises."""
    import torch

    from asr_study_b200.core.models import CTCModel
    m = CTCModel.__new__(CTCModel)
    m.device, m.history = torch.device("cpu"), {}
    calls = []

    def stats(x):
        k = float(x[0][0, 0, 0])
        calls.append(k)
        return torch.tensor([k + 1.0, k, 0.0, 0.5])

    monkeypatch.setattr(m, "_train_stats", stats, raising=False)
    checks = []
    monkeypatch.setattr(m, "check_status", lambda: checks.append(1), raising=False)   # the per-epoch watchdog read-back
    log = []
    ends = []

    class CB(object):
        def set_model(self, model):
            self.model = model

        def on_epoch_end(self, epoch, logs):
            ends.append((epoch, dict(logs)))

    g = _gen(4, log)
    hist = m.fit_generator(g, samples_per_epoch=12, nb_epoch=3, callbacks=[CB()], verbose=0, max_q_size=2, nb_worker=1)
    assert calls == [float(k) for k in range(9)] and log == list(range(9))
    assert [e for e, _ in ends] == [0, 1, 2] and len(checks) == 3
    np.testing.assert_allclose(hist["ctc_loss"], [1.0, 4.0, 7.0])          # batch means 0,1,2 | 3,4,5 | 6,7,8
    np.testing.assert_allclose(hist["loss"], [2.0, 5.0, 8.0])
    np.testing.assert_allclose(hist["decoder_ler"], [0.5, 0.5, 0.5])
    assert next(g)[0][0][0, 0, 0] == 9.0

    def boom(x):
        raise RuntimeError("step failed")

    monkeypatch.setattr(m, "_train_stats", boom, raising=False)
    import threading
    before = threading.active_count()
    with pytest.raises(RuntimeError, match="step failed"):
        m.fit_generator(_gen(4, []), samples_per_epoch=400, nb_epoch=1, verbose=0, max_q_size=2, nb_worker=1)
    time.sleep(0.5)
    assert threading.active_count() <= before                            # the worker was stopped
