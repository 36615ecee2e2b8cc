"""ctc_model(inputs, output) on symbolic layer records (core/models.py:31-52, README.md:96-108 of the reference): the
graph a user writes with Input / GaussianNoise / Dropout / TimeDistributed(Dense) / Bidirectional(LSTM) / merge lowers to
the engine's ModelSpec; what the hot path does not build is rejected loudly.  CPU: the CTCModel constructor (which
needs a GPU) is replaced by a recorder."""
import pytest

from asr_study_b200.core import models
from asr_study_b200.core.layers import (LSTM, Bidirectional, Dense, Dropout, GaussianNoise, Input, TimeDistributed, l2,
                                        merge, recurrent)


@pytest.fixture
def record(monkeypatch):
    made = []

    class Rec(object):
        def __init__(self, spec, **kw):
            self.spec, self.kw = spec, kw
            made.append(self)

    monkeypatch.setattr(models, "CTCModel", Rec)
    return made


def test_readme_custom_model_recipe(record):
    def custom_model(num_features=26, num_hiddens=100, num_classes=28):
        x = Input(name='inputs', shape=(None, num_features))
        o = x
        o = Bidirectional(LSTM(num_hiddens, return_sequences=True, consume_less='gpu'))(o)
        o = TimeDistributed(Dense(num_classes))(o)
        return models.ctc_model(x, o)

    m = custom_model()
    s = m.spec
    assert (s.num_features, s.num_hiddens, s.num_layers, s.num_classes) == (26, 100, 1, 28)
    assert s.dropout == 0.0 and s.weight_decay == 0.0 and not s.general and m.kw["input_std_noise"] == 0.0


def test_factories_lower_like_the_reference_topologies(record):
    g = models.graves2006()
    assert (g.spec.num_features, g.spec.num_hiddens, g.spec.num_layers, g.spec.name) == (26, 100, 1, "graves2006")
    assert g.kw["input_std_noise"] == 0.6
    b = models.brsmv1()                                   # core/models.py:217-220 defaults
    s = b.spec
    assert (s.num_features, s.num_hiddens, s.num_layers, s.num_classes) == (39, 256, 5, 28)
    assert s.dropout == 0.2 and s.weight_decay == 1e-4 and s.residual is None and not s.general and s.name == "brsmv1"
    c2 = models.brsmv1(num_features=26, num_hiddens=512, num_layers=3).spec
    assert (c2.num_features, c2.num_hiddens, c2.num_layers) == (26, 512, 3)
    full = models.brsmv1(num_features=13, num_hiddens=64, num_layers=2, dropout=0.1, zoneout=0.2, input_dropout=True,
                         input_std_noise=0.3, residual="sum", layer_norm=[1.0, 0.0], mi=[1.0, 0.5, 0.5])
    s = full.spec
    assert s.residual == "sum" and s.input_dropout and s.zoneout == 0.2 and s.layer_norm == (1.0, 0.0)
    assert s.mi == (1.0, 0.5, 0.5) and s.proj_width == 128 and s.general and full.kw["input_std_noise"] == 0.3
    e = models.eyben()
    assert e.spec.layer_hiddens == (120, 27) and e.spec.input_dense == 78 and e.spec.num_features == 39
    e2 = models.eyben(num_hiddens=[0, 50, 0])
    assert e2.spec.num_layers == 1 and e2.spec.input_dense is None and e2.spec.num_hiddens == 50
    with pytest.raises(NotImplementedError):
        models.maas()


def test_unsupported_graphs_are_rejected_loudly(record):
    x = Input(name="inputs", shape=(None, 26))
    with pytest.raises(NotImplementedError):              # a unidirectional layer is not on the hot path
        LSTM(32, return_sequences=True)(x)
    with pytest.raises(NotImplementedError):
        Bidirectional(LSTM(32))                          # return_sequences=False
    with pytest.raises(NotImplementedError):
        Bidirectional(LSTM(32, return_sequences=True), merge_mode="sum")
    with pytest.raises(NotImplementedError):
        models.brsmv1(residual="concat")
    with pytest.raises(NotImplementedError):
        LSTM(32, activation="relu")
    with pytest.raises(NotImplementedError):
        recurrent(32, model="gru")
    with pytest.raises(ValueError):
        recurrent(32, model="nope")
    o = Bidirectional(LSTM(32, return_sequences=True))(x)
    with pytest.raises(NotImplementedError):              # no logits layer
        models.ctc_model(x, o)
    with pytest.raises(NotImplementedError):              # weight decay not tied
        models.ctc_model(x, TimeDistributed(Dense(28, W_regularizer=l2(1e-3)))(o))
    y = Bidirectional(LSTM(32, return_sequences=True, dropout_W=0.1, dropout_U=0.3))(x)
    with pytest.raises(NotImplementedError):
        models.ctc_model(x, TimeDistributed(Dense(28))(y))
    other = Input(name="other", shape=(None, 26))
    with pytest.raises(NotImplementedError):              # logits do not descend from `inputs`
        models.ctc_model(other, TimeDistributed(Dense(28))(o))
    with pytest.raises(NotImplementedError):              # residual without the 2H projection
        r = merge([o, o], mode="sum")
        models.ctc_model(x, TimeDistributed(Dense(28))(r))
    with pytest.raises(TypeError):
        models.ctc_model(26, [LSTM(10), 28])


def test_dropout_and_noise_records_pass_through(record):
    x = Input(name="inputs", shape=(None, 20))
    o = GaussianNoise(0.25)(x)
    o = Dropout(0.2)(o)
    o = Bidirectional(recurrent(64, model="lstm", dropout=0.2, return_sequences=True))(o)
    m = models.ctc_model(x, TimeDistributed(Dense(30))(o), beam_width=50)
    assert m.spec.input_dropout and m.spec.dropout == 0.2 and m.spec.num_classes == 30 and m.spec.general
    assert m.kw["input_std_noise"] == 0.25 and m.kw["beam_width"] == 50
