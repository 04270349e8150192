# -*- coding: utf-8 -*-
"""Dump golden vectors of the projection loop FROM THE UNMODIFIED REFERENCE (TensorFlow 1.x, Python 2.7).

This container has neither TF1 nor Python 2, so parity of this repo is pinned only to its own CPU oracle
("parity unpinned" in DESIGN.md).  Run this script on a machine that has the reference's environment
(README.md:46: Python 2.7, tensorflow 1.7, keras 2.1.5) to produce the missing pin:

    cd /path/to/defensegan                                   # the reference checkout (kabkabm/defensegan)
    python /path/to/repo/tools/make_tf_golden.py --golden /path/to/repo/tests/golden --out /path/to/repo/tests/golden/tf

It is written in the Python-2/TF-1 dialect of the reference on purpose (no f-strings, no type hints).  For every
fixture case of tests/golden/ (same seeded weights, images and z0 as the oracle fixtures) it

  1. builds the reference model class (models/gan.py:649-765) from the reference's own yml,
  2. overwrites the Generator variables with the fixture's seeded weights (restated below exactly as
     oracle/defensegan_oracle.py:init_generator_weights draws them; the SHA-256 is checked against the fixture),
  3. builds DefenseGANBase.reconstruct (models/gan.py:333-449) with z_init_val = the fixture's z0 (the reference's own
     hook, gan.py:395-397), runs tf.local_variables_initializer() + the op exactly like utils/gan_defense.py:119,146,
  4. writes <out>/<case>.npz {rec_tf, loss_min_tf, tf_version} and saves the generator with the reference's saver
     naming (`GAN.model-<step>` + `checkpoint`, base_model.py:383-395) under <out>/<case>_ckpt/ so that
     defensegan_b200/tf_bundle.py can be checked against a file TensorFlow itself wrote.

tests/test_tf_golden.py picks the files up when they exist (and is skipped while they do not).
"""
from __future__ import print_function

import argparse
import collections
import hashlib
import os
import sys

import numpy as np

CASES = ['mnist_c1', 'mnist_ragged_bias', 'celeba_small']


def deconv_channels(arch, net_dim):
    if arch == 'mnist':
        return [('Generator.2', 4 * net_dim, 2 * net_dim), ('Generator.3', 2 * net_dim, net_dim),
                ('Generator.5', net_dim, 1)]
    return [('Generator.2', 4 * net_dim, 2 * net_dim), ('Generator.3', 2 * net_dim, net_dim),
            ('Generator.5', net_dim, net_dim), ('Generator.6', net_dim, 3)]


def seeded_weights(arch, random_bias, seed=11241990, latent_dim=128, net_dim=64):
    """Same draws, same order as oracle/defensegan_oracle.py:init_generator_weights (no BatchNorm)."""
    rs = np.random.RandomState(seed)
    w = collections.OrderedDict()

    def uniform(stdev, size):
        return rs.uniform(low=-stdev * np.sqrt(3), high=stdev * np.sqrt(3), size=size).astype('float32')

    n_feat = 4 * 4 * 4 * net_dim
    w['Generator.Input/Generator.Input.W'] = uniform(np.sqrt(2.0 / (latent_dim + n_feat)), (latent_dim, n_feat))
    w['Generator.Input/Generator.Input.b'] = np.zeros((n_feat,), dtype='float32')
    for name, c_in, c_out in deconv_channels(arch, net_dim):
        fan_in = c_in * 25 / 4.0
        fan_out = c_out * 25
        w['%s/%s.Filters' % (name, name)] = uniform(np.sqrt(4.0 / (fan_in + fan_out)), (5, 5, c_out, c_in))
        w['%s/%s.Biases' % (name, name)] = np.zeros((c_out,), dtype='float32')
    if random_bias:
        for key in list(w.keys()):
            if key.endswith('.b') or key.endswith('.Biases'):
                w[key] = (0.1 * rs.standard_normal(w[key].shape)).astype('float32')
    return w


def digest(w):
    h = hashlib.sha256()
    for k, v in w.items():
        h.update(k.encode('utf-8'))
        h.update(np.ascontiguousarray(v).tobytes())
    return h.hexdigest()


def run_case(case, golden_dir, out_dir):
    import tensorflow as tf
    from utils.config import load_config
    from models.gan import MnistDefenseGAN, CelebADefenseGAN

    g = np.load(os.path.join(golden_dir, case + '.npz'))
    arch = str(g['arch'])
    B, R, L, lr = int(g['B']), int(g['R']), int(g['L']), float(g['lr'])
    weights = seeded_weights(arch, bool(int(g['random_bias'])))
    assert digest(weights) == str(g['weights_sha256']), 'weight restatement differs from the fixture'

    tf.reset_default_graph()
    tf.set_random_seed(11241990)
    cfg = load_config('experiments/cfgs/gans/%s.yml' % arch)
    cls = MnistDefenseGAN if arch == 'mnist' else CelebADefenseGAN
    # batch_size must be a multiple of rec_rr (models/gan.py:101-104); the graph is static in it
    gan = cls(cfg=cfg, test_mode=True, verbose=False, batch_size=B * R)
    gan.rec_rr, gan.rec_iters, gan.rec_lr = R, L, lr
    sess = gan.sess
    sess.run(tf.global_variables_initializer())

    # overwrite the Generator variables with the seeded weights, by tflib.param name (tflib/__init__.py:7-33)
    by_name = dict((v.name.split(':')[0], v) for v in gan.generator_vars)
    for name, val in weights.items():
        assert name in by_name, (name, sorted(by_name))
        sess.run(tf.assign(by_name[name], val))

    images_pl = tf.placeholder(tf.float32, shape=[B] + list(g['images'].shape[1:]))
    z0 = tf.constant(g['z0'].astype('float32'))
    rec_op = gan.reconstruct(images_pl, batch_size=B, back_prop=False, reconstructor_id=0, z_init_val=z0)
    sess.run(tf.local_variables_initializer())                      # utils/gan_defense.py:119
    rec = sess.run(rec_op, feed_dict={images_pl: g['images']})      # utils/gan_defense.py:146
    loss_min = ((rec - g['images']) ** 2).reshape(B, -1).mean(axis=1)

    if not os.path.isdir(out_dir):
        os.makedirs(out_dir)
    np.savez_compressed(os.path.join(out_dir, case + '.npz'), rec_tf=rec.astype('float32'),
                        loss_min_tf=loss_min.astype('float32'), tf_version=tf.__version__)
    ckpt_dir = os.path.join(out_dir, case + '_ckpt')
    if not os.path.isdir(ckpt_dir):
        os.makedirs(ckpt_dir)
    saver = tf.train.Saver(var_list=gan.generator_vars)
    saver.save(sess, os.path.join(ckpt_dir, gan.model_save_name), global_step=1)
    print('[%s] wrote rec %s, loss_min[:4] %s, checkpoint in %s' % (case, rec.shape, loss_min[:4], ckpt_dir))
    print('[%s] vs fixture: max|rec_tf - rec32| = %.3g, max|loss_min_tf - loss_min32| = %.3g' % (
        case, float(np.abs(rec - g['rec32']).max()), float(np.abs(loss_min - g['loss_min32']).max())))
    gan.close_session()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--reference', default='.', help='checkout of kabkabm/defensegan (the current directory by default)')
    ap.add_argument('--golden', required=True, help='tests/golden of this repository')
    ap.add_argument('--out', required=True, help='where to write <case>.npz and <case>_ckpt/ (tests/golden/tf)')
    ap.add_argument('--cases', nargs='*', default=CASES)
    args = ap.parse_args()
    golden, out = os.path.abspath(args.golden), os.path.abspath(args.out)
    os.chdir(args.reference)                       # utils/config.py reads experiments/cfgs/key_doc.yml relatively
    sys.path.insert(0, os.getcwd())
    for case in args.cases:
        run_case(case, golden, out)


if __name__ == '__main__':
    main()
