"""Classifiers and the reconstruction layer: mirror of reference utils/network_builder.py.

* `ReconstructionLayer` / `add_rec_model` (reference :179-183, 239-271) put gan.reconstruct() in front of a classifier
  (layer 0, reconstructor_id 123).
* `MLP` + the layer classes `Conv2D / Linear / ReLU / Dropout / Softmax / Flatten` (reference :139-331) and the model
  zoo `model_a ... model_f, model_q, model_y, model_z` (reference :333-521) in PyTorch, NHWC in and out like the
  reference (TF `SAME` / `VALID` padding rules, kernels drawn column-normalised as in :203-208, 222-227).  They are the
  downstream classifiers of the "chosen reconstructions and downstream classifier accuracy must match" clause of
  BASELINE.json's north_star: the hot path ends where their input begins.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


class ReconstructionLayer(torch.nn.Module):
    """reference utils/network_builder.py:239-271: fprop(x) = gan.reconstruct(x, batch_size,
    back_prop, reconstructor_id=123, z_init_val)."""

    def __init__(self, model, input_shape, batch_size, z_init_val=None, back_prop=True):
        super().__init__()
        self.rec_model = [model]          # not a sub-module: the GAN is frozen, non-torch state
        self.input_shape = input_shape
        self.output_shape = input_shape
        self.batch_size = batch_size
        self.z_init_val = z_init_val
        self.back_prop = back_prop

    def set_input_shape(self, shape):
        self.input_shape = shape
        self.output_shape = shape

    def get_output_shape(self):
        return self.output_shape

    def fprop(self, x):
        x = x.reshape((x.shape[0],) + tuple(self.input_shape[1:]))
        return self.rec_model[0].reconstruct(x, batch_size=None, back_prop=self.back_prop, reconstructor_id=123,
                                             z_init_val=self.z_init_val)

    forward = fprop


class DefendedModel(torch.nn.Module):
    """classifier(reconstruct(x)) - what `MLP.add_rec_model` builds (reference :179-183)."""

    def __init__(self, classifier, gan, input_shape, batch_size=None, z_init_val=None):
        super().__init__()
        self.rec_layer = ReconstructionLayer(gan, input_shape, batch_size, z_init_val=z_init_val)
        self.classifier = classifier

    def forward(self, x):
        return self.classifier(self.rec_layer(x))

    get_probs = forward


def add_rec_model(classifier, gan, input_shape, batch_size=None, z_init_val=None):
    return DefendedModel(classifier, gan, input_shape, batch_size=batch_size, z_init_val=z_init_val)


# ---------------------------------------------------------------------------------------------------------
# classifier zoo (reference utils/network_builder.py:139-521)
# ---------------------------------------------------------------------------------------------------------
class Layer(torch.nn.Module):
    """reference :186-188.  Shapes follow the reference convention [batch, H, W, C] / [batch, dim] (batch = None)."""

    def set_input_shape(self, shape):
        self.input_shape = list(shape)
        self.output_shape = list(shape)

    def get_output_shape(self):
        return self.output_shape

    def fprop(self, x):
        raise NotImplementedError

    def forward(self, x):
        return self.fprop(x)


class Linear(Layer):
    """x W + b; W ~ N(0,1) with every column scaled to unit L2 norm (reference :191-209)."""

    def __init__(self, num_hid):
        super().__init__()
        self.num_hid = int(num_hid)

    def set_input_shape(self, input_shape):
        batch_size, dim = input_shape
        self.input_shape = [batch_size, dim]
        self.output_shape = [batch_size, self.num_hid]
        init = torch.randn(dim, self.num_hid)
        init = init / torch.sqrt(1e-7 + (init ** 2).sum(dim=0, keepdim=True))
        self.W = torch.nn.Parameter(init)
        self.b = torch.nn.Parameter(torch.zeros(self.num_hid))

    def fprop(self, x):
        return x @ self.W + self.b


class Conv2D(Layer):
    """tf.nn.conv2d(x, kernels, strides, padding) + b on NHWC tensors; kernels (kh, kw, Cin, Cout) ~ N(0,1) normalised
    over (kh, kw, Cin) (reference :212-237).  `SAME` pads like TensorFlow: total = max((ceil(H/s) - 1) s + k - H, 0),
    the smaller half first."""

    def __init__(self, output_channels, kernel_shape, strides, padding):
        super().__init__()
        self.output_channels = int(output_channels)
        self.kernel_shape = tuple(int(k) for k in kernel_shape)
        self.strides = tuple(int(s) for s in strides)
        self.padding = str(padding).upper()
        assert self.padding in ("SAME", "VALID")

    def _pads(self, rows, cols):
        if self.padding == "VALID":
            return (0, 0, 0, 0)
        out = []
        for size, k, s in ((cols, self.kernel_shape[1], self.strides[1]), (rows, self.kernel_shape[0], self.strides[0])):
            total = max((int(math.ceil(size / float(s))) - 1) * s + k - size, 0)
            out += [total // 2, total - total // 2]
        return tuple(out)             # (left, right, top, bottom) for F.pad on NCHW

    def set_input_shape(self, input_shape):
        batch_size, rows, cols, input_channels = input_shape
        kh, kw = self.kernel_shape
        init = torch.randn(kh, kw, input_channels, self.output_channels)
        init = init / torch.sqrt(1e-7 + (init ** 2).sum(dim=(0, 1, 2)))
        self.kernels = torch.nn.Parameter(init)                      # TF layout (kh, kw, Cin, Cout)
        self.b = torch.nn.Parameter(torch.zeros(self.output_channels))
        self.input_shape = list(input_shape)
        with torch.no_grad():
            dummy = self.fprop(torch.zeros(1, rows, cols, input_channels))
        self.output_shape = [batch_size] + list(dummy.shape[1:])

    def fprop(self, x):
        xn = x.permute(0, 3, 1, 2)
        pads = self._pads(x.shape[1], x.shape[2])
        if any(pads):
            xn = F.pad(xn, pads)
        y = F.conv2d(xn, self.kernels.permute(3, 2, 0, 1), self.b, stride=self.strides)
        return y.permute(0, 2, 3, 1)


class ReLU(Layer):
    def fprop(self, x):
        return torch.relu(x)


class Dropout(Layer):
    """tf.nn.dropout(x, keep_prob=prob) while K.learning_phase() (= module.training here), identity otherwise
    (reference :289-302: the constructor argument is TF1's KEEP probability)."""

    def __init__(self, prob):
        super().__init__()
        self.prob = float(prob)

    def fprop(self, x):
        return F.dropout(x, p=1.0 - self.prob, training=self.training)


class Softmax(Layer):
    def fprop(self, x):
        return torch.softmax(x, dim=-1)


class Flatten(Layer):
    def set_input_shape(self, shape):
        self.input_shape = list(shape)
        width = 1
        for f in shape[1:]:
            width *= f
        self.output_width = width
        self.output_shape = [None, width]

    def fprop(self, x):
        return x.reshape(-1, self.output_width)


class MLP(torch.nn.Module):
    """reference :139-183: a list of layers with named states; `fprop` returns {layer name: output}; the last
    Softmax is 'probs' and the layer before it 'logits'."""

    def __init__(self, layers, input_shape, rec_model=None):
        super().__init__()
        self.layer_names = []
        self.input_shape = list(input_shape)
        if isinstance(layers[-1], Softmax):
            layers[-1].name = "probs"
            layers[-2].name = "logits"
        else:
            layers[-1].name = "logits"
        shape = list(input_shape)
        for i, layer in enumerate(layers):
            self.layer_names.append(getattr(layer, "name", layer.__class__.__name__ + str(i)))
            layer.set_input_shape(shape)
            shape = layer.get_output_shape()
        self.layers = torch.nn.ModuleList(layers)
        self._rec_layer = None
        self.rec_model = rec_model

    def get_layer_names(self):
        return (["reconstruction"] if self._rec_layer is not None else []) + list(self.layer_names)

    def fprop(self, x, set_ref=False, no_rec=False):
        states = {}
        if self._rec_layer is not None and not no_rec:
            x = self._rec_layer.fprop(x)
            states["reconstruction"] = x
        for name, layer in zip(self.layer_names, self.layers):
            if set_ref:
                layer.ref = x
            x = layer.fprop(x)
            states[name] = x
        return states

    def get_logits(self, x, **kw):
        return self.fprop(x, **kw)["logits"]

    def get_probs(self, x, **kw):
        st = self.fprop(x, **kw)
        return st["probs"] if "probs" in st else torch.softmax(st["logits"], dim=-1)

    def forward(self, x):
        """Logits (what the losses and arg-max consumers of this package take)."""
        return self.get_logits(x)

    def add_rec_model(self, model, z_init, batch_size):
        """reference :179-183: reconstruction becomes layer 0."""
        self._rec_layer = ReconstructionLayer(model, self.input_shape, batch_size, z_init_val=z_init)
        self._rec_layer.set_input_shape(self.input_shape)


def model_f(nb_filters=64, nb_classes=10, input_shape=(None, 28, 28, 1), rec_model=None):
    return MLP([Conv2D(nb_filters, (8, 8), (2, 2), "SAME"), ReLU(), Conv2D(nb_filters * 2, (6, 6), (2, 2), "VALID"), ReLU(),
                Conv2D(nb_filters * 2, (5, 5), (1, 1), "VALID"), ReLU(), Flatten(), Linear(nb_classes), Softmax()],
               input_shape, rec_model=rec_model)


def model_e(input_shape=(None, 28, 28, 1), nb_classes=10):
    return MLP([Flatten(), Linear(200), ReLU(), Linear(200), ReLU(), Linear(nb_classes), Softmax()], input_shape)


def model_d(input_shape=(None, 28, 28, 1), nb_classes=10):
    return MLP([Flatten(), Linear(200), ReLU(), Dropout(0.5), Linear(200), ReLU(), Linear(nb_classes), Softmax()], input_shape)


def model_b(nb_filters=64, nb_classes=10, input_shape=(None, 28, 28, 1), rec_model=None):
    return MLP([Dropout(0.2), Conv2D(nb_filters, (8, 8), (2, 2), "SAME"), ReLU(), Conv2D(nb_filters * 2, (6, 6), (2, 2), "VALID"),
                ReLU(), Conv2D(nb_filters * 2, (5, 5), (1, 1), "VALID"), ReLU(), Dropout(0.5), Flatten(), Linear(nb_classes),
                Softmax()], input_shape, rec_model=rec_model)


def model_a(nb_filters=64, nb_classes=10, input_shape=(None, 28, 28, 1), rec_model=None):
    return MLP([Conv2D(nb_filters, (5, 5), (1, 1), "SAME"), ReLU(), Conv2D(nb_filters, (5, 5), (2, 2), "VALID"), ReLU(),
                Flatten(), Dropout(0.25), Linear(128), ReLU(), Dropout(0.5), Linear(nb_classes), Softmax()],
               input_shape, rec_model=rec_model)


def model_c(nb_filters=64, nb_classes=10, input_shape=(None, 28, 28, 1), rec_model=None):
    return MLP([Conv2D(nb_filters * 2, (3, 3), (1, 1), "SAME"), ReLU(), Conv2D(nb_filters, (5, 5), (2, 2), "VALID"), ReLU(),
                Flatten(), Dropout(0.25), Linear(128), ReLU(), Dropout(0.5), Linear(nb_classes), Softmax()],
               input_shape, rec_model=rec_model)


def _conv_stack(specs):
    out = []
    for ch, k, s, pad in specs:
        out += [Conv2D(ch, (k, k), (s, s), pad), ReLU()]
    return out


def model_y(nb_filters=64, nb_classes=10, input_shape=(None, 28, 28, 1), rec_model=None):
    f = nb_filters
    return MLP(_conv_stack([(f, 3, 1, "SAME"), (f, 3, 2, "VALID"), (2 * f, 3, 2, "VALID"), (2 * f, 3, 2, "VALID")]) +
               [Flatten(), Linear(256), ReLU(), Dropout(0.5), Linear(256), ReLU(), Dropout(0.5), Linear(nb_classes), Softmax()],
               input_shape, rec_model=rec_model)


def model_q(nb_filters=32, nb_classes=10, input_shape=(None, 28, 28, 1), rec_model=None):
    f = nb_filters
    return MLP(_conv_stack([(f, 3, 1, "SAME"), (f, 3, 2, "VALID"), (2 * f, 3, 1, "VALID"), (2 * f, 3, 2, "VALID")]) +
               [Flatten(), Linear(256), ReLU(), Dropout(0.5), Linear(256), ReLU(), Dropout(0.5), Linear(nb_classes), Softmax()],
               input_shape, rec_model=rec_model)


def model_z(nb_filters=32, nb_classes=10, input_shape=(None, 28, 28, 1), rec_model=None):
    f = nb_filters
    return MLP(_conv_stack([(f, 3, 1, "SAME"), (f, 3, 2, "VALID"), (2 * f, 3, 1, "VALID"), (2 * f, 3, 2, "VALID"),
                            (4 * f, 3, 1, "VALID"), (4 * f, 3, 2, "VALID")]) +
               [Flatten(), Linear(600), ReLU(), Dropout(0.5), Linear(600), ReLU(), Dropout(0.5), Linear(nb_classes), Softmax()],
               input_shape, rec_model=rec_model)


# reference blackbox.py:372-386 / whitebox.py:97-111 pick the classifier by letter
model_dict = {"A": model_a, "B": model_b, "C": model_c, "D": model_d, "E": model_e, "F": model_f, "Q": model_q,
              "Y": model_y, "Z": model_z}
