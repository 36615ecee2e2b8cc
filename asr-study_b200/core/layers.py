"""Recurrent-layer descriptors with the reference's constructor surface
(core/layers.py:356-516).  A layer here is a *configuration record*: the arithmetic of
LSTM.step (core/layers.py:432-469) lives in the persistent CUDA kernels; the model factories
collect these records into the engine's ModelSpec.

Built: the default step plus variational dropout (dropout_W / dropout_U) on the tensor-core engines; zoneout, layer
normalisation and multiplicative integration on the general-cell engine (csrc/lstm_cell.cu).  Other activations are
rejected loudly, never silently ignored.
"""


class LSTM(object):
    def __init__(self, output_dim, zoneout_h=0., zoneout_c=0., layer_norm=None, mi=None, return_sequences=True,
                 consume_less="gpu", activation="tanh", inner_activation="hard_sigmoid", W_regularizer=None,
                 U_regularizer=None, dropout_W=0., dropout_U=0., go_backwards=False, **kwargs):
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
        if not return_sequences:
            raise NotImplementedError("return_sequences=False is not used by any reference topology")
        self.output_dim = int(output_dim)
        self.zoneout_h = self.zoneout_c = float(zoneout_h)
        self.layer_norm = None if layer_norm is None else tuple(float(v) for v in layer_norm)
        self.mi = None if mi is None else tuple(float(v) for v in mi)
        self.dropout_W, self.dropout_U = float(dropout_W), float(dropout_U)
        self.W_regularizer, self.U_regularizer = W_regularizer, U_regularizer
        self.consume_less = "gpu"

    def get_config(self):
        return {"output_dim": self.output_dim, "layer_norm": self.layer_norm, "mi": self.mi,
                "zoneout_h": self.zoneout_h, "zoneout_c": self.zoneout_c}


def recurrent(output_dim, model="keras_lstm", activation="tanh", regularizer=None, dropout=0., **kwargs):
    """core/layers.py:482-516 — only the LSTM families are built."""
    if model in ("keras_lstm", "lstm"):
        return LSTM(output_dim, activation=activation, W_regularizer=regularizer, U_regularizer=regularizer,
                    dropout_W=dropout, dropout_U=dropout, **kwargs)
    if model in ("rnn", "gru", "rhn"):
        raise NotImplementedError("model %s is outside the BiLSTM hot path" % model)
    raise ValueError("model %s was not recognized" % model)
