"""
On-disk formats either side of the hot path, without the MySQL hop of the
reference:

* ``users.dat`` of the citeulike datasets -- one line per user, ``<count> id id ...``
  (/root/reference/util/data_parser.py:200-219 reads it into the ``articles_users`` table,
  :101-115 turns the table into the dense 0/1 ratings matrix).  ``read_users_dat`` goes straight to CSR
  with the same id conventions: citeulike-a ids are 0-based in the file (the importer adds 1, the matrix
  subtracts it again), citeulike-t ids are 1-based;
* the top-recommendations dump of ``AbstractRecommender.dump_recommendations``
  (/root/reference/lib/abstract_recommender.py:148-171): one line per user, ``<count> id id ...`` with
  1-based ids of the recommendations scoring above 1e-6, ascending score (the order
  ``TopRecommendations`` stores them in); ``<count>`` alone for a user without any.
"""
import numpy
import scipy.sparse as sp

ID_BASE = {'citeulike-a': 0, 'citeulike-t': 1}       # data_parser.py:211-214
SCORE_FLOOR = 1e-6                                   # abstract_recommender.py:160


def read_users_dat(path, dataset='citeulike-a', n_items=None):
    """users x items CSR (float64 ones) of a ``users.dat`` file.

    :param str dataset: 'citeulike-a' (0-based article ids) or 'citeulike-t' (1-based), data_parser.py:211-214.
    :param int n_items: number of articles (the reference counts the rows of its ``articles`` table); default:
        the largest id seen + 1.
    A repeated (user, article) pair counts once (the table's primary key / the matrix assignment ``= 1``)."""
    if dataset not in ID_BASE:
        raise ValueError("dataset must be one of %s, got %r" % (sorted(ID_BASE), dataset))
    base = ID_BASE[dataset]
    indptr, indices = [0], []
    with open(path) as f:
        for line in f:
            fields = line.replace("\n", "").split(" ")
            count = int(fields[0])
            row = sorted(set(int(x) - base for x in fields[1:count + 1]))
            if row and row[0] < 0:
                raise ValueError("article id below %d in a %s file" % (base, dataset))
            indices.extend(row)
            indptr.append(len(indices))
    width = (max(indices) + 1 if indices else 0) if n_items is None else int(n_items)
    if indices and max(indices) >= width:
        raise ValueError("article id %d does not fit %d articles" % (max(indices) + base, width))
    return sp.csr_matrix((numpy.ones(len(indices)), numpy.asarray(indices, dtype=numpy.int32),
                          numpy.asarray(indptr, dtype=numpy.int64)), shape=(len(indptr) - 1, width))


def write_users_dat(path, ratings, dataset='citeulike-a'):
    """Inverse of ``read_users_dat`` (dense or sparse ratings; non-zeros are the library)."""
    base = ID_BASE[dataset]
    csr = sp.csr_matrix(ratings)
    csr.eliminate_zeros()
    csr.sort_indices()
    with open(path, "w") as f:
        for user in range(csr.shape[0]):
            ids = csr.indices[csr.indptr[user]:csr.indptr[user + 1]] + base
            f.write(' '.join([str(len(ids))] + [str(int(i)) for i in ids]) + '\n')


def recommendation_lines(topk_idx, topk_val):
    """Dump lines from the device's top-N arrays (users x N, DESCENDING score, -1 padded -- the layout
    hyprec_eval_topk returns): ascending score, ids + 1, scores <= 1e-6 dropped."""
    lines = []
    for ids, vals in zip(numpy.asarray(topk_idx), numpy.asarray(topk_val)):
        keep = (ids >= 0) & (vals > SCORE_FLOOR)
        chosen = [str(int(i) + 1) for i in ids[keep][::-1]]
        lines.append('%d %s\n' % (len(chosen), ' '.join(chosen)) if chosen else '%d\n' % len(chosen))
    return lines


def read_recommendations(path):
    """Lists of 0-based item ids per user from a dump written by ``dump_recommendations``."""
    out = []
    with open(path) as f:
        for line in f:
            fields = line.split()
            out.append([int(x) - 1 for x in fields[1:int(fields[0]) + 1]])
    return out
