"""`ReconstructionLayer` / `add_rec_model`: mirror of reference utils/network_builder.py:179-183,
239-271 - puts gan.reconstruct() in front of a classifier (layer 0, reconstructor_id 123)."""
from __future__ import annotations

import torch


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
