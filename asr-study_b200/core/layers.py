"""Layer records with the constructor surface the reference's model code uses (core/layers.py:356-516 and the Keras-1
names core/models.py imports: Input, GaussianNoise, Dropout, Dense, TimeDistributed, Bidirectional, merge, l2).

A layer here is a *configuration record* and calling it on a symbolic tensor extends a tiny graph; nothing is computed.
``core.models.ctc_model(inputs, output)`` walks that graph from ``output`` back to ``inputs`` and lowers it to the
engine's ModelSpec — the arithmetic of LSTM.step (core/layers.py:432-469) lives in the persistent CUDA kernels.  This is
what makes the README's custom-model recipe (README.md:96-108) work unchanged:

    x = Input(name='inputs', shape=(None, num_features))
    o = Bidirectional(LSTM(num_hiddens, return_sequences=True, consume_less='gpu'))(x)
    o = TimeDistributed(Dense(num_classes))(o)
    model = ctc_model(x, o)

Anything the hot path does not build (other activations, merge modes that change the width, SimpleRNN / GRU / RHN,
unidirectional stacks) is rejected loudly, never silently ignored.
"""


class Tensor(object):
    """symbolic tensor: the record that produced it, its parent tensors and its Keras shape (None = batch / time)."""

    def __init__(self, producer, parents, shape, name=None):
        self.producer, self.parents, self.shape, self.name = producer, list(parents), tuple(shape), name

    _keras_shape = property(lambda self: self.shape)


def Input(shape=None, name=None, dtype="float32", sparse=False, **kwargs):
    return Tensor(None, [], (None,) + tuple(shape), name=name)


class _Record(object):
    def __call__(self, x):
        return Tensor(self, [x], self.output_shape(x.shape))

    def output_shape(self, shape):
        return shape


class l2(object):
    """keras.regularizers.l2 record (core/models.py:263-264, 279)."""

    def __init__(self, l=0.01):
        self.l2 = float(l)


def _decay(reg):
    if reg is None:
        return 0.0
    if isinstance(reg, l2):
        return reg.l2
    if isinstance(reg, (int, float)):
        return float(reg)
    raise NotImplementedError("only l2 weight regularisers are built (core/models.py:263-264)")


class GaussianNoise(_Record):
    def __init__(self, sigma, **kwargs):
        self.sigma = float(sigma or 0.0)


class Dropout(_Record):
    def __init__(self, p, **kwargs):
        if not 0.0 <= float(p) < 1.0:
            raise ValueError("dropout must be in [0, 1)")
        self.p = float(p)


class Dense(_Record):
    def __init__(self, output_dim, activation=None, W_regularizer=None, b_regularizer=None, **kwargs):
        if activation not in (None, "linear"):
            raise NotImplementedError("Dense layers of the hot path are linear (core/models.py:71, 278)")
        if b_regularizer is not None:
            raise NotImplementedError("bias regularisers are not used by any reference topology")
        self.output_dim, self.weight_decay = int(output_dim), _decay(W_regularizer)

    def output_shape(self, shape):
        return tuple(shape[:-1]) + (self.output_dim,)


class TimeDistributed(_Record):
    def __init__(self, layer, **kwargs):
        if not isinstance(layer, Dense):
            raise NotImplementedError("TimeDistributed wraps Dense on the hot path")
        self.layer = layer

    def output_shape(self, shape):
        return self.layer.output_shape(shape)


class LSTM(_Record):
    def __init__(self, output_dim, zoneout_h=0., zoneout_c=0., layer_norm=None, mi=None, return_sequences=False,
                 consume_less="gpu", activation="tanh", inner_activation="hard_sigmoid", W_regularizer=None,
                 U_regularizer=None, b_regularizer=None, dropout_W=0., dropout_U=0., go_backwards=False, **kwargs):
        if float(zoneout_h) != float(zoneout_c):
            raise NotImplementedError("zoneout_h and zoneout_c are tied (core/models.py:267-268)")
        if not 0.0 <= float(zoneout_h) < 1.0:
            raise ValueError("zoneout must be in [0, 1)")
        if layer_norm is not None and len(layer_norm) != 2:
            raise ValueError("layer_norm = [gain_init, bias_init] (core/layers.py:408)")
        if mi is not None and len(mi) != 3:
            raise ValueError("mi = [alpha_init, beta1_init, beta2_init] (core/layers.py:392)")
        if activation != "tanh" or inner_activation != "hard_sigmoid":
            raise NotImplementedError("only tanh / hard_sigmoid (the Keras-1 defaults) are built")
        if not (0.0 <= dropout_W < 1.0 and 0.0 <= dropout_U < 1.0):
            raise ValueError("dropout must be in [0, 1)")
        if b_regularizer is not None:
            raise NotImplementedError("bias regularisers are not used by any reference topology")
        self.output_dim = int(output_dim)
        self.zoneout_h = self.zoneout_c = float(zoneout_h)
        self.layer_norm = None if layer_norm is None else tuple(float(v) for v in layer_norm)
        self.mi = None if mi is None else tuple(float(v) for v in mi)
        self.dropout_W, self.dropout_U = float(dropout_W), float(dropout_U)
        self.W_regularizer, self.U_regularizer = W_regularizer, U_regularizer
        self.return_sequences, self.go_backwards = bool(return_sequences), bool(go_backwards)
        self.consume_less = "gpu"                       # core/layers.py:383-386 forces it

    def __call__(self, x):
        raise NotImplementedError("the hot path stacks Bidirectional(LSTM(...)); a unidirectional LSTM layer is not built")

    def output_shape(self, shape):
        return tuple(shape[:-1]) + (self.output_dim,)

    def get_config(self):
        return {"output_dim": self.output_dim, "layer_norm": self.layer_norm, "mi": self.mi,
                "zoneout_h": self.zoneout_h, "zoneout_c": self.zoneout_c, "dropout_W": self.dropout_W,
                "dropout_U": self.dropout_U, "return_sequences": self.return_sequences,
                "go_backwards": self.go_backwards}


class Bidirectional(_Record):
    def __init__(self, layer, merge_mode="concat", **kwargs):
        if not isinstance(layer, LSTM):
            raise NotImplementedError("Bidirectional wraps core.layers.LSTM on the hot path")
        if merge_mode != "concat":
            raise NotImplementedError("Bidirectional(merge_mode=%r): the reference uses the Keras default 'concat'" % merge_mode)
        if not layer.return_sequences:
            raise NotImplementedError("return_sequences=False is not used by any reference topology")
        self.layer = layer

    def output_shape(self, shape):
        return tuple(shape[:-1]) + (2 * self.layer.output_dim,)


class _Merge(_Record):
    def __init__(self, mode):
        self.mode = mode


def merge(inputs, mode="sum", **kwargs):
    """keras.layers.merge as core/models.py:273-274 uses it: merge([new_o, o], mode=residual)."""
    if len(inputs) != 2:
        raise NotImplementedError("merge of two tensors (the residual connection) is what the hot path builds")
    if mode != "sum":
        raise NotImplementedError("merge mode %r: only 'sum' keeps the layer width the next Bidirectional expects" % mode)
    if inputs[0].shape[-1] != inputs[1].shape[-1]:
        raise ValueError("merge(mode='sum') needs equal widths, got %r and %r" % (inputs[0].shape, inputs[1].shape))
    return Tensor(_Merge(mode), list(inputs), inputs[0].shape)


def recurrent(output_dim, model="keras_lstm", activation="tanh", regularizer=None, dropout=0., **kwargs):
    """core/layers.py:482-516 — only the LSTM families are built."""
    if model in ("keras_lstm", "lstm"):
        return LSTM(output_dim, activation=activation, W_regularizer=regularizer, U_regularizer=regularizer,
                    dropout_W=dropout, dropout_U=dropout, **kwargs)
    if model in ("rnn", "gru", "rhn"):
        raise NotImplementedError("model %s is outside the BiLSTM hot path" % model)
    raise ValueError("model %s was not recognized" % model)
