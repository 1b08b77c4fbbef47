"""GMM initialisation of the VaDE latent space from embeddings — the reference's ``VaDEPT.initialize_gmm_from_data``
(``deepof/clustering/models_new.py:1907-1947``) calls scikit-learn's ``GaussianMixture`` (an un-vendored third-party
dependency of this path; scikit-learn 1.9 in this image) with these exact arguments and stores ``means_`` and
``log(covariances_)``.  Off the step path: runs once per fit on at most 10 000 embeddings."""
from typing import Tuple

import numpy as np


def gmm_from_embeddings(embeddings: np.ndarray, n_components: int, reg_covar: float = 1e-4) -> Tuple[np.ndarray, np.ndarray]:
    """(means [K, D], log-variances [K, D]).  Uses numpy's global RNG for the k-means initialisation exactly like the
    reference (``random_state=None``; ``train_deepof_model`` seeds it with ``np.random.seed(random_seed)``)."""
    from sklearn.mixture import GaussianMixture
    gmm = GaussianMixture(n_components=int(n_components), covariance_type="diag", reg_covar=reg_covar).fit(np.asarray(embeddings))
    return gmm.means_, np.log(gmm.covariances_)
