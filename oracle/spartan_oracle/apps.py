"""ORACLE (test infrastructure only) -- k-means iteration and strip SpMV.

PINNED by vectors produced by running the reference's own mapper functions: oracle/ref_apps/make_app_vectors.py takes
the definitions of kmeans_map2_dist_mapper / kmeans_count_mapper / kmeans_center_mapper (k_means_.py:61-97) and
dot_map2_mapper (dot.py:195-217) out of the reference's files, executes them on seeded inputs and commits inputs +
outputs as tests/golden/app_vectors.json; tests/test_oracle_reference_vectors.py checks every function below against
them (labels, counts and partial products bit for bit).  The reference's own tests for these two paths assert nothing
(tests/test_kmeans.py:16-21, tests/test_pagerank.py, tests/test_sparse.py), which is why the
vectors had to be generated.

k-means: spartan/examples/sklearn/cluster/k_means_.py:61-97 (mappers) and :130-160 (driver, 'map2').
SpMV:    spartan/expr/dot.py:213-217 (`tocsr().dot(dense)` per column strip) + np.add merge.
"""
import numpy as np
import scipy.sparse
from scipy.spatial.distance import cdist


def kmeans_dist_mapper(points, centers):
  return np.argmin(cdist(points, centers), axis=1)                 # k_means_.py:61-66


def kmeans_count_mapper(labels, centers_count):
  return np.bincount(labels.astype(np.int64), minlength=centers_count)   # k_means_.py:69-72


def kmeans_center_mapper(points, labels, centers_count):
  new_centers = np.zeros((centers_count, points.shape[1]))         # k_means_.py:75-97
  for i in range(centers_count):
    new_centers[i] = points[labels == i].sum(axis=0)
  return new_centers


def kmeans_fit(X, centers, n_iter, tile_rows, seed=0):
  """k_means_.py:130-160 with the per-tile results SUMMED across row tiles (the evident intent; the
  reference's reducer-less map2 targets overwrite) and a seeded re-seed of empty
  clusters."""
  rng = np.random.RandomState(seed)
  n, d = X.shape
  k = centers.shape[0]
  centers = centers.astype(np.float64)
  labels = np.zeros(n, dtype=np.int64)
  for _ in range(n_iter):
    counts = np.zeros(k)
    sums = np.zeros((k, d))
    for r0 in range(0, n, tile_rows):
      pts = X[r0:r0 + tile_rows]
      lab = kmeans_dist_mapper(pts, centers)
      labels[r0:r0 + tile_rows] = lab
      counts += kmeans_count_mapper(lab, k)
      sums += kmeans_center_mapper(pts, lab, k)
    z = counts == 0
    if np.any(z):
      counts[z] = 1
      sums[z, :] = rng.randn(int(np.count_nonzero(z)), d)
    centers = sums / counts.reshape(k, 1)
  return centers, labels


def spmv_strips(matrix, x, strip_width):
  """Per column strip `tiles[0].tocsr().dot(tiles[1])`, partials merged with np.add (dot.py:213-217)."""
  matrix = scipy.sparse.csc_matrix(matrix)
  n_rows, n_cols = matrix.shape
  y = np.zeros(n_rows, dtype=np.float32)
  xv = np.asarray(x).reshape(-1)
  for c0 in range(0, n_cols, strip_width):
    c1 = min(n_cols, c0 + strip_width)
    y = np.add(y, matrix[:, c0:c1].tocsr().dot(xv[c0:c1]).astype(np.float32))
  return y


def make_weights(n, outlinks, seed):
  """tests/benchmark_pagerank.py:11-24 with a seeded generator: n*outlinks random (dest, source) pairs."""
  rng = np.random.default_rng(seed)
  num_out = n * outlinks
  source = rng.integers(0, n, num_out)
  dest = rng.integers(0, n, num_out)
  value = rng.random(num_out, dtype=np.float32)
  return scipy.sparse.coo_matrix((value, (source, dest)), shape=(n, n))
