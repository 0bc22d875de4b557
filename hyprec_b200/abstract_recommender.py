"""
Mirror of the reference's recommender contract (/root/reference/lib/abstract_recommender.py:10-204):
same method names, arguments and return values.  The dense-matrix methods stay available for API
compatibility; subclasses that own device state override the heavy ones
(``get_evaluation_report``, ``recommend_items``) with C-ABI calls.
"""
import os

import numpy

from hyprec_b200.top_recommendations import top_k_stream_order


class AbstractRecommender(object):
    def __init__(self, initializer, evaluator, hyperparameters, options, **flags):
        raise NotImplementedError("Can't initialize this class")

    def __repr__(self):
        return self.__class__.__name__

    def set_options(self, options):
        """abstract_recommender.py:21-41: n_iterations, k_folds; any option ``x`` overrides an
        existing attribute ``_x``."""
        self.n_iter = options['n_iterations']
        self.k_folds = options['k_folds']
        self.splitting_method = 'kfold'
        self._split_type = 'user'
        self.evaluator.set_kfolds(self.k_folds)
        if self.k_folds == 1:
            self.splitting_method = 'naive'
        self.options = options.copy()
        for option, value in options.items():
            if hasattr(self, '_' + option):
                setattr(self, '_' + option, value)

    def set_hyperparameters(self, hyperparameters):
        raise NotImplementedError("Can't call this method")

    def predict(self, user, item):
        raise NotImplementedError("Can't call this method")

    def train_one_fold(self):
        raise NotImplementedError("Can't call this method")

    def train(self):
        if self.splitting_method == 'naive':
            self.set_data(*self.evaluator.naive_split(self._split_type))
            return self.train_one_fold()
        self.fold_test_indices = self.evaluator.get_kfold_indices()
        return self.train_k_fold()

    def set_data(self, train_data, test_data):
        self.predictions = None
        self.train_data = train_data
        self.test_data = test_data

    def train_k_fold(self):
        all_errors = []
        for current_k in range(self.k_folds):
            self.set_data(*self.evaluator.get_fold(current_k, self.fold_test_indices))
            self.hyperparameters['fold'] = current_k
            self.train_one_fold()
            all_errors.append(self.get_evaluation_report())
        return numpy.mean(all_errors, axis=0)

    def get_predictions(self):
        raise NotImplementedError("Can't call this method")

    def get_ratings(self):
        return self.ratings

    def rounded_predictions(self):
        """abstract_recommender.py:173-187: 1 where score >= row average else 0; the average is a
        sequential float64 sum / n (numpy.cumsum adds in the same order as Python's ``sum``)."""
        predictions = self.get_predictions()
        avg = numpy.cumsum(predictions, axis=1)[:, -1] / predictions.shape[1]
        return (predictions >= avg[:, None]).astype(predictions.dtype)

    def recommend_items(self, user_id, num_recommendations=10):
        """abstract_recommender.py:189-204: all items streamed in index order; (id, score) pairs in
        ASCENDING score order, as TopRecommendations stores them."""
        scores = self.get_predictions()[user_id]
        ids, vals = top_k_stream_order(numpy.arange(len(scores)), scores, num_recommendations)
        return zip(reversed(ids), reversed(vals))

    def dump_recommendations(self, num_recommendations=10):
        """abstract_recommender.py:148-171: one line per user, ``<count> id id ...`` with 1-based ids
        of the recommendations scoring above 1e-6."""
        lines = []
        for user in range(self.ratings.shape[0]):
            keep = [str(idx + 1) for idx, val in self.recommend_items(user, num_recommendations) if val > 1e-6]
            lines.append('%d %s\n' % (len(keep), ' '.join(keep)) if keep else '%d\n' % len(keep))
        base_dir = os.path.dirname(os.path.realpath(__file__))
        path = os.path.join(os.path.dirname(base_dir), 'matrices/%s' % self.results_file_name)
        with open(path, "w") as f:
            f.writelines(lines)
        if self._verbose:
            print("dumped top recommendations to %s" % path)

    def get_evaluation_report(self):
        """abstract_recommender.py:100-131 on dense matrices (generic path for subclasses without
        device state)."""
        predictions = self.get_predictions()
        rounded_predictions = self.rounded_predictions()
        test_sum = self.test_data.sum()
        train_sum = self.train_data.sum()
        ev = self.evaluator
        ev.load_top_recommendations(200, predictions, self.test_data, self.hyperparameters['fold'])
        train_recall = ev.calculate_recall(self.train_data, rounded_predictions)
        test_recall = ev.calculate_recall(self.test_data, rounded_predictions)
        recall_at_x = ev.recall_at_x(200, predictions, self.test_data, rounded_predictions)
        total = self.ratings.sum() if hasattr(self.ratings, 'todense') else sum(sum(self.ratings))
        ratio = sum(sum(rounded_predictions)) / total
        mrr_at_five = ev.calculate_mrr(5, predictions, self.test_data, rounded_predictions)
        ndcg_at_five = ev.calculate_ndcg(5, predictions, self.test_data, rounded_predictions)
        mrr_at_ten = ev.calculate_mrr(10, predictions, self.test_data, rounded_predictions)
        ndcg_at_ten = ev.calculate_ndcg(10, predictions, self.test_data, rounded_predictions)
        rmse = ev.get_rmse(predictions, self.train_data)
        report = (test_sum, train_sum, rmse, train_recall, test_recall, recall_at_x, ratio, mrr_at_five,
                  ndcg_at_five, mrr_at_ten, ndcg_at_ten)
        if self._verbose:
            print(REPORT_FORMAT.format(*report))
        return report


REPORT_FORMAT = ('Test sum {:.2f}, Train sum {:.2f}, Final error {:.5f}, train recall {:.5f}, '
                 'test recall {:.5f}, recall@200 {:.5f}, ratio {:.5f}, mrr@5 {:.5f}, '
                 'ndcg@5 {:.5f}, mrr@10 {:.5f}, ndcg@10 {:.5f}')
