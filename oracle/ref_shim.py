"""Minimal Keras-1.2.2 / TF-1.3 look-alike on torch.float64 (CPU) that lets the REFERENCE'S OWN modules
(/root/reference/core/{layers,layers_utils,initializers,ctc_utils,models}.py, preprocessing/audio.py) execute
verbatim under Python 3.12, so that golden vectors come from the reference's code and not from a restatement.

TEST INFRASTRUCTURE ONLY (oracle/): used by oracle/make_golden_lstm.py and oracle/make_golden.py in the build
container, where /root/reference is mounted.  Nothing here travels into the product path, and nothing in tests/
needs /root/reference at run time (they read the committed fixtures under tests/golden/).

What runs verbatim (reference code):  LSTM.__init__/build/step/get_config (core/layers.py:366-479),
layer_normalization / zoneout / multiplicative_integration (core/layers_utils.py:16-51), k_init
(core/initializers.py), the topologies graves2006 / eyben / brsmv1 and ctc_model (core/models.py:31-103,217-281),
ctc_lambda_func / decode (core/ctc_utils.py:8-70), Feature/FBank/MFCC/LogFbank (preprocessing/audio.py).

What this file restates (un-vendored third-party code, pinned in the reference's msc.yaml: keras==1.2.2,
tensorflow==1.3.0) — each item is the published behaviour of that version, kept as small as the reference's call
sites need:
  keras.layers.LSTM (recurrent.py of 1.2.2): constructor defaults, consume_less='gpu' build (W [D,4H], U [H,4H],
      b = [0 | forget_bias_init | 0 | 0]), get_constants (B_U / B_W: four K.dropout(ones) draws each in the train
      phase, the scalar 1 otherwise), call = K.rnn over time with go_backwards reversing the INPUT, zero initial states
  keras.layers.Bidirectional (wrappers.py): forward copy + from_config(get_config() with go_backwards flipped),
      the backward outputs reversed back, merge_mode='concat'
  TimeDistributed(Dense), GaussianNoise / Dropout (train phase only), merge(mode='sum'), Lambda, Input, Model
  K.dot / K.sqrt / K.dropout (keep mask / (1 - level), optional noise_shape) / K.in_train_phase / hard_sigmoid =
      clip(0.2 x + 0.5, 0, 1) / tf.nn.moments (population variance)
  tf.nn.ctc_loss -> torch.nn.functional.ctc_loss on log_softmax(logits) with blank = C - 1 (an INDEPENDENT CTC
      implementation, not oracle/ctc.py); tf.nn.ctc_greedy_decoder -> first-max argmax, merge repeats, drop blanks
Because tensors are torch.float64 with autograd, d(loss)/d(parameter) of the reference's own forward code is
available as well (TF's autodiff is replaced by torch's; both differentiate the same expression graph).
"""
from __future__ import annotations

import copy
import importlib.util
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
DT = torch.float64


class Ctx:
    """Evaluation context: learning phase, the random source of K.dropout / GaussianNoise and a log of every mask drawn."""

    def __init__(self):
        self.training = False
        self.rng = np.random.RandomState(0)
        self.dropout_log = []          # (tag, keep mask as float64 ndarray, level)
        self.noise_log = []

    def reset(self, training, seed):
        self.training, self.rng = training, np.random.RandomState(seed)
        self.dropout_log, self.noise_log = [], []


CTX = Ctx()


def _t(x):
    if isinstance(x, torch.Tensor):
        return x
    return torch.as_tensor(np.asarray(x), dtype=DT)


# ------------------------------------------------------------------ keras.backend
def _make_backend():
    K = types.ModuleType("keras.backend")
    K.dot = lambda a, b: torch.matmul(_t(a), _t(b))
    K.sqrt = lambda x: torch.sqrt(_t(x))
    K.square = lambda x: _t(x) ** 2
    K.sum = lambda x, axis=None: _t(x).sum() if axis is None else _t(x).sum(axis)
    K.floatx = lambda: "float64"
    K.cast_to_floatx = lambda x: float(x)
    K.zeros = lambda shape, **kw: torch.zeros(shape, dtype=DT)
    K.is_sparse = lambda x: False

    def variable(value, dtype=None, name=None):
        v = torch.tensor(np.asarray(value, dtype=np.float64), dtype=DT, requires_grad=True)
        v.keras_name = name
        return v

    K.variable = variable

    def dropout(x, level, noise_shape=None, seed=None):
        x = _t(x)
        shape = tuple(x.shape) if noise_shape is None else tuple(noise_shape)
        keep = (CTX.rng.uniform(size=shape) >= level).astype(np.float64)
        CTX.dropout_log.append((shape, keep, float(level)))
        return x * _t(keep) / (1.0 - level)

    K.dropout = dropout
    K.in_train_phase = lambda a, b: a if CTX.training else b
    K.concatenate = lambda xs, axis=-1: torch.cat([_t(x) for x in xs], dim=axis)
    K.reverse = lambda x, axes: torch.flip(_t(x), [axes] if isinstance(axes, int) else list(axes))
    K.ones_like = lambda x: torch.ones_like(_t(x))
    K.tanh = lambda x: torch.tanh(_t(x))
    K.hard_sigmoid = lambda x: torch.clamp(0.2 * _t(x) + 0.5, 0.0, 1.0)
    return K


# ------------------------------------------------------------------ tensorflow
def _make_tf():
    tf = types.ModuleType("tensorflow")
    tf.int32 = "int32"
    tf.nn = types.SimpleNamespace()

    def moments(x, axes, keep_dims=False):
        x = _t(x)
        mean = x.mean(dim=axes, keepdim=keep_dims)
        var = ((x - x.mean(dim=axes, keepdim=True)) ** 2).mean(dim=axes, keepdim=keep_dims)   # population variance
        return mean, var

    tf.nn.moments = moments
    tf.cast = lambda x, dtype: x.to(torch.int64) if isinstance(x, torch.Tensor) else np.asarray(x).astype(np.int64)
    tf.transpose = lambda x, perm: _t(x).permute(*perm)

    def ctc_loss(labels, inputs, sequence_length, preprocess_collapse_repeated=False, ctc_merge_repeated=True):
        """labels: list of int sequences (the SparseTensor of the reference); inputs: time-major logits [T, N, C]."""
        assert not preprocess_collapse_repeated and ctc_merge_repeated          # TF-1.3 defaults, the reference's call
        logits = _t(inputs)
        T, N, Cc = logits.shape
        lens = torch.as_tensor(np.asarray(sequence_length).reshape(-1), dtype=torch.int64)
        tl = torch.as_tensor([len(l) for l in labels], dtype=torch.int64)
        flat = torch.as_tensor(np.concatenate([np.asarray(l, np.int64) for l in labels]))
        lp = torch.log_softmax(logits, dim=-1)
        return torch.nn.functional.ctc_loss(lp, flat, lens, tl, blank=Cc - 1, reduction="none", zero_infinity=False)

    tf.nn.ctc_loss = ctc_loss

    def ctc_greedy_decoder(inputs, sequence_length, merge_repeated=True):
        x = _t(inputs).detach().numpy()
        T, N, Cc = x.shape
        out = []
        for n in range(N):
            prev, seq = -1, []
            for t in range(int(np.asarray(sequence_length)[n])):
                k = int(np.argmax(x[t, n]))                  # first maximum wins
                if k != Cc - 1 and not (merge_repeated and k == prev):
                    seq.append(k)
                prev = k
            out.append(seq)
        return [out], None

    tf.nn.ctc_greedy_decoder = ctc_greedy_decoder
    tf.sparse_tensor_to_dense = lambda x, default_value=-1: x
    return tf


# ------------------------------------------------------------------ symbolic graph (Input / Layer / Model)
class Sym:
    def __init__(self, fn, parents, shape, name=None):
        self.fn, self.parents, self._keras_shape, self.name = fn, parents, shape, name

    def eval(self, feeds, cache):
        if id(self) in cache:
            return cache[id(self)]
        if self.fn is None:
            v = feeds[self.name]
        else:
            v = self.fn(*[p.eval(feeds, cache) for p in self.parents])
        cache[id(self)] = v
        return v


class Layer:
    _uid = 0

    def __init__(self, **kwargs):
        Layer._uid += 1
        self.name = kwargs.get("name") or "%s_%d" % (self.__class__.__name__.lower(), Layer._uid)
        self.trainable = kwargs.get("trainable", True)
        self.built = False
        self.trainable_weights = []
        self.regularizers = []
        self.uses_learning_phase = False

    def add_weight(self, shape, initializer, name=None, regularizer=None, trainable=True):
        w = initializer(shape, name=name)
        if regularizer is not None:
            self.regularizers.append((regularizer, w))
        self.trainable_weights.append(w)
        return w

    def build(self, input_shape):
        self.built = True

    def get_output_shape_for(self, input_shape):
        return input_shape

    def __call__(self, x):
        shp = [s._keras_shape for s in x] if isinstance(x, (list, tuple)) else x._keras_shape
        if not self.built:
            self.build(shp)
            self.built = True
        parents = list(x) if isinstance(x, (list, tuple)) else [x]
        multi = isinstance(x, (list, tuple))
        fn = (lambda *vals: self.call(list(vals))) if multi else (lambda v: self.call(v))
        sym = Sym(fn, parents, self.get_output_shape_for(shp), name=self.name)
        sym.layer = self
        return sym

    def get_config(self):
        return {"name": self.name, "trainable": self.trainable}

    @classmethod
    def from_config(cls, config):
        return cls(**config)


def Input(name=None, shape=None, dtype="float32", sparse=False, **kw):
    return Sym(None, [], (None,) + tuple(shape), name=name)


class _Init:
    """keras.initializations of 1.2.2 on a seeded numpy stream (values are arbitrary test data; shapes / scales follow
    glorot_uniform, orthogonal(scale = 1.1), one, zero)."""
    rng = np.random.RandomState(0)

    @staticmethod
    def glorot_uniform(shape, name=None):
        s = np.sqrt(6.0 / (shape[0] + shape[1]))
        return KB.variable(_Init.rng.uniform(-s, s, size=shape), name=name)

    @staticmethod
    def orthogonal(shape, scale=1.1, name=None):
        a = _Init.rng.normal(0.0, 1.0, shape)
        u, _, v = np.linalg.svd(a, full_matrices=False)
        q = u if u.shape == tuple(shape) else v
        return KB.variable(scale * q.reshape(shape), name=name)

    @staticmethod
    def one(shape, name=None):
        return KB.variable(np.ones(shape), name=name)

    @staticmethod
    def zero(shape, name=None):
        return KB.variable(np.zeros(shape), name=name)

    @staticmethod
    def uniform(shape, scale=0.05, name=None):
        return KB.variable(_Init.rng.uniform(-scale, scale, size=shape), name=name)

    @staticmethod
    def get(x):
        return getattr(_Init, x) if isinstance(x, str) else x


class _Act:
    tanh = staticmethod(lambda x: torch.tanh(_t(x)))
    hard_sigmoid = staticmethod(lambda x: torch.clamp(0.2 * _t(x) + 0.5, 0.0, 1.0))
    sigmoid = staticmethod(lambda x: torch.sigmoid(_t(x)))
    linear = staticmethod(lambda x: x)
    relu = staticmethod(lambda x, alpha=0.0, max_value=None: torch.relu(_t(x)))

    @staticmethod
    def get(x):
        if x is None:
            return _Act.linear
        return getattr(_Act, x) if isinstance(x, str) else x


class L2:
    def __init__(self, l2=0.01):
        self.l2 = float(l2)

    def __call__(self, w):
        return self.l2 * (w ** 2).sum()

    def get_config(self):
        return {"name": "L1L2Regularizer", "l1": 0.0, "l2": self.l2}


def _reg_get(x):
    if x is None or isinstance(x, L2):
        return x
    if isinstance(x, dict):
        return L2(x.get("l2", 0.0))
    raise TypeError(x)


class KerasLSTM(Layer):
    """keras.layers.LSTM of Keras 1.2.2 (recurrent.py): the base class the reference's LSTM extends."""

    def __init__(self, output_dim, init="glorot_uniform", inner_init="orthogonal", forget_bias_init="one",
                 activation="tanh", inner_activation="hard_sigmoid", W_regularizer=None, U_regularizer=None,
                 b_regularizer=None, dropout_W=0.0, dropout_U=0.0, weights=None, return_sequences=False,
                 go_backwards=False, stateful=False, unroll=False, consume_less="cpu", input_dim=None,
                 input_length=None, **kwargs):
        super().__init__(**kwargs)
        self.output_dim = output_dim
        self._init_names = dict(init=init, inner_init=inner_init, forget_bias_init=forget_bias_init,
                                activation=activation, inner_activation=inner_activation)
        self.init, self.inner_init = _Init.get(init), _Init.get(inner_init)
        self.forget_bias_init = _Init.get(forget_bias_init)
        self.activation, self.inner_activation = _Act.get(activation), _Act.get(inner_activation)
        self.W_regularizer, self.U_regularizer = _reg_get(W_regularizer), _reg_get(U_regularizer)
        self.b_regularizer = _reg_get(b_regularizer)
        self.dropout_W, self.dropout_U = dropout_W, dropout_U
        if self.dropout_W or self.dropout_U:
            self.uses_learning_phase = True
        self.return_sequences, self.go_backwards = return_sequences, go_backwards
        self.stateful, self.unroll, self.consume_less = stateful, unroll, consume_less
        self.input_dim, self.input_length = input_dim, input_length

    def build(self, input_shape):
        self.input_dim = input_shape[2]
        H = self.output_dim
        assert self.consume_less == "gpu"         # the reference forces it (core/layers.py:383-386)
        self.W = self.add_weight((self.input_dim, 4 * H), initializer=self.init, name="%s_W" % self.name,
                                 regularizer=self.W_regularizer)
        self.U = self.add_weight((H, 4 * H), initializer=self.inner_init, name="%s_U" % self.name,
                                 regularizer=self.U_regularizer)

        def b_reg(shape, name=None):
            fb = self.forget_bias_init((H,)).detach().numpy()
            return KB.variable(np.hstack((np.zeros(H), fb, np.zeros(H), np.zeros(H))), name=name)

        self.b = self.add_weight((4 * H,), initializer=b_reg, name="%s_b" % self.name, regularizer=self.b_regularizer)
        self.built = True

    def get_output_shape_for(self, input_shape):
        return (input_shape[0], input_shape[1], self.output_dim) if self.return_sequences else (input_shape[0], self.output_dim)

    def get_constants(self, x):
        consts = []
        for level, width in ((self.dropout_U, self.output_dim), (self.dropout_W, self.input_dim)):
            if 0 < level < 1:
                ones = torch.ones(x.shape[0], width, dtype=DT)
                consts.append([KB.in_train_phase(KB.dropout(ones, level), ones) for _ in range(4)])
            else:
                consts.append([1.0 for _ in range(4)])
        return consts          # [B_U, B_W]

    def call(self, x, mask=None):
        x = _t(x)
        N, T, _ = x.shape
        states = [torch.zeros(N, self.output_dim, dtype=DT), torch.zeros(N, self.output_dim, dtype=DT)]
        constants = self.get_constants(x)
        order = range(T - 1, -1, -1) if self.go_backwards else range(T)      # K.rnn reverses the input
        outs = []
        for t in order:
            out, states = self.step(x[:, t], states + constants)
            outs.append(out)
        y = torch.stack(outs, dim=1)             # in processing order, like K.rnn
        return y if self.return_sequences else outs[-1]

    def get_config(self):
        cfg = dict(output_dim=self.output_dim, W_regularizer=self.W_regularizer.get_config() if self.W_regularizer else None,
                   U_regularizer=self.U_regularizer.get_config() if self.U_regularizer else None,
                   b_regularizer=self.b_regularizer.get_config() if self.b_regularizer else None,
                   dropout_W=self.dropout_W, dropout_U=self.dropout_U, return_sequences=self.return_sequences,
                   go_backwards=self.go_backwards, stateful=self.stateful, unroll=self.unroll,
                   consume_less=self.consume_less, input_dim=self.input_dim, input_length=self.input_length)
        cfg.update(self._init_names)
        return dict(list(super().get_config().items()) + list(cfg.items()))


class Bidirectional(Layer):
    def __init__(self, layer, merge_mode="concat", **kwargs):
        super().__init__(**kwargs)
        self.forward_layer = copy.copy(layer)
        config = layer.get_config()
        config["go_backwards"] = not config["go_backwards"]
        self.backward_layer = layer.__class__.from_config(config)
        self.forward_layer.name = "forward_" + self.forward_layer.name
        self.backward_layer.name = "backward_" + self.backward_layer.name
        self.merge_mode = merge_mode
        self.return_sequences = layer.return_sequences

    def build(self, input_shape):
        for l in (self.forward_layer, self.backward_layer):
            l.trainable_weights, l.regularizers = [], []
            l.build(input_shape)

    def get_output_shape_for(self, input_shape):
        s = self.forward_layer.get_output_shape_for(input_shape)
        return s[:-1] + (2 * s[-1],) if self.merge_mode == "concat" else s

    def call(self, x, mask=None):
        y = self.forward_layer.call(x, mask)
        y_rev = self.backward_layer.call(x, mask)
        if self.return_sequences:
            y_rev = KB.reverse(y_rev, 1)
        assert self.merge_mode == "concat"
        return KB.concatenate([y, y_rev])


class Dense(Layer):
    def __init__(self, output_dim, init="glorot_uniform", activation=None, W_regularizer=None, **kwargs):
        super().__init__(**kwargs)
        self.output_dim, self.init = output_dim, _Init.get(init)
        self.activation, self.W_regularizer = _Act.get(activation), _reg_get(W_regularizer)

    def build(self, input_shape):
        self.W = self.add_weight((input_shape[-1], self.output_dim), initializer=self.init, name=self.name + "_W",
                                 regularizer=self.W_regularizer)
        self.b = self.add_weight((self.output_dim,), initializer=_Init.zero, name=self.name + "_b")

    def get_output_shape_for(self, input_shape):
        return tuple(input_shape[:-1]) + (self.output_dim,)

    def call(self, x, mask=None):
        return self.activation(KB.dot(x, self.W) + self.b)


class TimeDistributed(Layer):
    def __init__(self, layer, **kwargs):
        super().__init__(**kwargs)
        self.layer = layer

    def build(self, input_shape):
        self.layer.build((input_shape[0],) + tuple(input_shape[2:]))

    def get_output_shape_for(self, input_shape):
        return tuple(input_shape[:2]) + self.layer.get_output_shape_for((input_shape[0],) + tuple(input_shape[2:]))[1:]

    def call(self, x, mask=None):
        return self.layer.call(_t(x))            # Dense broadcasts over [N, T, D]


class GaussianNoise(Layer):
    def __init__(self, sigma, **kwargs):
        super().__init__(**kwargs)
        self.sigma = sigma

    def call(self, x, mask=None):
        x = _t(x)
        if CTX.training and self.sigma:
            noise = CTX.rng.normal(0.0, self.sigma, size=tuple(x.shape))
            CTX.noise_log.append(noise)
            return x + _t(noise)
        return x


class Dropout(Layer):
    def __init__(self, p, **kwargs):
        super().__init__(**kwargs)
        self.p = p

    def call(self, x, mask=None):
        if 0.0 < self.p < 1.0 and CTX.training:
            return KB.dropout(x, self.p)
        return _t(x)


class Lambda(Layer):
    def __init__(self, function, output_shape=None, arguments=None, **kwargs):
        super().__init__(**kwargs)
        self.function, self.arguments = function, arguments or {}

    def get_output_shape_for(self, input_shape):
        return None

    def call(self, x, mask=None):
        return self.function(x, **self.arguments)


def merge(inputs, mode="sum", **kw):
    assert mode == "sum", "only the width-preserving merge is exercised"
    return Sym(lambda *v: sum(_t(a) for a in v), list(inputs), inputs[0]._keras_shape, name="merge")


class Model:
    def __init__(self, input, output, name=None):
        self.inputs, self.outputs = list(input), list(output)

    def run(self, feeds, training=False, seed=0, want=()):
        """-> outputs (and the values of the symbolic tensors in `want`, evaluated in the same pass)."""
        CTX.reset(training, seed)
        cache = {}
        outs = [o.eval(feeds, cache) for o in self.outputs]
        return (outs, [w.eval(feeds, cache) for w in want]) if want else outs

    def layers(self):
        seen, out = set(), []

        def walk(s):
            if id(s) in seen:
                return
            seen.add(id(s))
            for p in s.parents:
                walk(p)
            out.append(s)

        for o in self.outputs:
            walk(o)
        return out


KB = _make_backend()
TF = _make_tf()


def install():
    """Put the look-alike modules into sys.modules and load the reference's core package by path."""
    keras = types.ModuleType("keras")
    keras.backend = KB
    mods = {"keras": keras, "keras.backend": KB, "tensorflow": TF}

    def sub(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        mods[name] = m
        return m

    inits = sub("keras.initializations", get=_Init.get, one=_Init.one, zero=_Init.zero, uniform=_Init.uniform,
                glorot_uniform=_Init.glorot_uniform, orthogonal=_Init.orthogonal)
    acts = sub("keras.activations", get=_Act.get, relu=_Act.relu, tanh=_Act.tanh, hard_sigmoid=_Act.hard_sigmoid)
    regs = sub("keras.regularizers", l2=L2, l1=None, l1l2=None)
    layers = sub("keras.layers", LSTM=KerasLSTM, GRU=None, SimpleRNN=None, Input=Input, GaussianNoise=GaussianNoise,
                 TimeDistributed=TimeDistributed, Dense=Dense, Masking=None, Bidirectional=Bidirectional,
                 Lambda=Lambda, Dropout=Dropout, merge=merge)
    layers.recurrent = sub("keras.layers.recurrent", Recurrent=Layer)
    keras.engine = sub("keras.engine", Layer=Layer, InputSpec=object)
    keras.models = sub("keras.models", Model=Model)
    keras.initializations, keras.activations, keras.regularizers, keras.layers = inits, acts, regs, layers
    sys.modules.update(mods)

    def load(name, path, pkg=False):
        spec = importlib.util.spec_from_file_location(name, path, submodule_search_locations=[path.rsplit("/", 1)[0]] if pkg else None)
        m = importlib.util.module_from_spec(spec)
        sys.modules[name] = m
        spec.loader.exec_module(m)
        return m

    # the packages' own __init__.py use py2 implicit-relative imports: stand-in package objects, real modules inside
    for pkg in ("core", "utils"):
        p = types.ModuleType(pkg)
        p.__path__ = [f"{REF}/{pkg}"]
        sys.modules[pkg] = p
    load("utils.hparams", f"{REF}/utils/hparams.py")
    out = {}
    for m in ("initializers", "layers_utils", "layers", "ctc_utils", "models"):
        out[m] = load(f"core.{m}", f"{REF}/core/{m}.py")
        setattr(sys.modules["core"], m, out[m])
    return out


def load_reference_audio():
    """preprocessing/audio.py itself (not a replay): py2 names and the scipy / librosa symbols it touches at import
    time are supplied; the ndarray branch of Feature.__call__ (audio.py:60-61) never reaches librosa."""
    import builtins
    import scipy.signal
    import scipy.signal.windows
    if not hasattr(scipy.signal, "hamming"):
        scipy.signal.hamming = scipy.signal.windows.hamming            # scipy 0.19's name (audio.py:182)
    sys.modules.setdefault("librosa", types.ModuleType("librosa"))
    for k, v in (("unicode", str), ("xrange", range)):
        if not hasattr(builtins, k):
            setattr(builtins, k, v)
    pkg = types.ModuleType("preprocessing")
    pkg.__path__ = [f"{REF}/preprocessing"]
    sys.modules["preprocessing"] = pkg
    for name in ("audio_utils", "audio"):
        spec = importlib.util.spec_from_file_location(f"preprocessing.{name}", f"{REF}/preprocessing/{name}.py")
        m = importlib.util.module_from_spec(spec)
        sys.modules[f"preprocessing.{name}"] = m
        spec.loader.exec_module(m)
        setattr(pkg, name, m)
    return sys.modules["preprocessing.audio"]
