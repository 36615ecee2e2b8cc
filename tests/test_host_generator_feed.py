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
