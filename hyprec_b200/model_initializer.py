"""
Host-side mirror of /root/reference/util/model_initializer.py (ModelInitializer, :8-103): factor
matrices are pickled with ``ndarray.dump`` to ``<folder>/<sorted key-value pairs><name>.dat``
(:71-103).  Differences: the folder is configurable and created on demand, and loading passes
``allow_pickle=True`` (the reference's bare ``numpy.load`` raises on these files under
numpy >= 1.16.3).
"""
import os

import numpy


class ModelInitializer(object):
    def __init__(self, config, n_iterations, verbose=False, folder=None):
        self.folder = folder if folder is not None else os.path.join(
            os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'matrices')
        self.set_config(config, n_iterations)
        self._v = verbose

    def set_config(self, config, n_iterations):
        self.config = config
        self.config['n_iterations'] = n_iterations

    def save_matrix(self, matrix, matrix_name, shape=None):
        path = self._create_path(matrix_name, matrix.shape if shape is None else shape)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        matrix.dump(path)
        if self._v:
            print("dumped to %s" % path)

    def load_matrix(self, config, matrix_name, matrix_shape):
        path = self._create_path(matrix_name, matrix_shape, config.copy())
        try:
            res = (True, numpy.load(path, allow_pickle=True))
            if self._v:
                print("loaded from %s" % path)
            return res
        except FileNotFoundError:
            if self._v:
                print("File not found, %s will initialize randomly" % path)
            return (False, numpy.random.random(matrix_shape))

    def _generate_file_name(self, config, matrix_name):
        pairs = ['%s-%s' % (key, str(config[key]).replace('.', '_')) for key in sorted(config)]
        return ','.join(pairs) + matrix_name

    def _create_path(self, matrix_name, matrix_shape, config=None):
        if config is None:
            config = self.config
        if 'n_iterations' not in config.keys():
            config['n_iterations'] = self.config['n_iterations']
        config['n_factors'] = matrix_shape[1]
        config['n_rows'] = matrix_shape[0]
        return os.path.join(self.folder, self._generate_file_name(config, matrix_name) + '.dat')
