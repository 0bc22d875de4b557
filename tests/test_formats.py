"""users.dat ingestion and the top-recommendations dump against literal
restatements of /root/reference/util/data_parser.py:101-115, 200-219 and
/root/reference/lib/abstract_recommender.py:148-171, 189-204."""
import numpy
import pytest

from hyprec_b200 import formats
from hyprec_b200.top_recommendations import TopRecommendations


def literal_ratings_matrix(lines, dataset, n_users, n_articles):
    """import_users (article_id = id + 1 for citeulike-a, id for citeulike-t; user ids from 1) followed by
    get_ratings_matrix (matrix[user_id - 1][article_id - 1] = 1)."""
    table = []
    for user_id, line in enumerate(lines, start=1):
        splitted = line.replace("\n", "").split(" ")
        for i in range(1, int(splitted[0]) + 1):
            table.append((user_id, int(splitted[i]) + 1 if dataset == 'citeulike-a' else int(splitted[i])))
    matrix = [[0] * n_articles for _ in range(n_users)]
    for user_id, article_id in table:
        matrix[user_id - 1][article_id - 1] = 1
    return numpy.array(matrix)


@pytest.mark.parametrize("dataset", ["citeulike-a", "citeulike-t"])
def test_users_dat_reader_matches_the_importer(tmp_path, dataset):
    rs = numpy.random.RandomState(11)
    base = formats.ID_BASE[dataset]
    n_users, n_articles = 17, 40
    lines = []
    for user in range(n_users):
        count = 0 if user == 5 else int(rs.randint(1, 12))
        ids = list(rs.choice(n_articles, size=count, replace=False) + base)
        if user == 2:
            ids.append(ids[0])                                   # a repeated pair counts once
        lines.append(' '.join([str(len(ids))] + [str(int(i)) for i in ids]) + '\n')
    path = tmp_path / "users.dat"
    path.write_text(''.join(lines))
    csr = formats.read_users_dat(str(path), dataset, n_items=n_articles)
    assert csr.shape == (n_users, n_articles) and csr.indices.dtype == numpy.int32
    assert numpy.array_equal(csr.toarray(), literal_ratings_matrix(lines, dataset, n_users, n_articles))
    assert csr.has_sorted_indices or numpy.all(numpy.diff(csr.indptr) <= 1)
    # round trip
    formats.write_users_dat(str(tmp_path / "again.dat"), csr, dataset)
    again = formats.read_users_dat(str(tmp_path / "again.dat"), dataset, n_items=n_articles)
    assert (again != csr).nnz == 0
    with pytest.raises(ValueError):
        formats.read_users_dat(str(path), dataset, n_items=3)
    with pytest.raises(ValueError):
        formats.read_users_dat(str(path), "movielens")


def literal_dump_lines(predictions, n):
    lines = []
    for user in range(predictions.shape[0]):
        top = TopRecommendations(n)
        for i in range(predictions.shape[1]):
            top.insert(i, predictions[user][i])
        recs = [str(idx + 1) for idx, val in zip(top.get_indices(), top.get_values()) if val > 1e-6]
        lines.append('%d %s\n' % (len(recs), ' '.join(recs)) if recs else '%d\n' % len(recs))
    return lines


def test_recommendation_lines_match_the_literal_dump(tmp_path):
    rs = numpy.random.RandomState(3)
    m, n, cap = 9, 30, 7
    predictions = numpy.round(rs.rand(m, n), 1)                   # plenty of ties
    predictions[4, :] = 0.0                                       # nothing above the floor
    predictions[6, 5:] = 1e-7                                     # fewer than cap above the floor
    idx = numpy.full((m, cap), -1, dtype=numpy.int32)
    val = numpy.zeros((m, cap))
    for user in range(m):                                         # what the device returns: descending, tie order
        top = TopRecommendations(cap)
        for i in range(n):
            top.insert(i, predictions[user][i])
        ids = list(reversed(top.get_indices()))
        idx[user, :len(ids)] = ids
        val[user, :len(ids)] = list(reversed(top.get_values()))
    lines = formats.recommendation_lines(idx, val)
    assert lines == literal_dump_lines(predictions, cap)
    assert lines[4] == '0\n'
    path = tmp_path / "top.dat"
    path.write_text(''.join(lines))
    back = formats.read_recommendations(str(path))
    assert back[4] == [] and len(back) == m
    assert back[0] == [int(x) - 1 for x in lines[0].split()[1:]]
