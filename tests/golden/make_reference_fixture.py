"""The data recipe of the reference's own test-suite (tests/conftest.py:22-93: sklearn
make_multilabel_classification(seed 2024, 100 000 x 50 features, 25 labels) + one logistic regression per
label), subsampled to 4 000 validation and 4 000 test rows and stored as float32 / uint8, so that
tests/test_reference_scenarios.py can replay the reference's tests on the GPU box (no sklearn fit there).

    python tests/golden/make_reference_fixture.py
"""
import os

import numpy as np
from sklearn.datasets import make_multilabel_classification
from sklearn.linear_model import LogisticRegression
from sklearn.model_selection import train_test_split
from sklearn.multioutput import MultiOutputClassifier

HERE = os.path.dirname(os.path.abspath(__file__))
seed = 2024
x, y = make_multilabel_classification(n_samples=100000, n_features=50, n_classes=25, n_labels=3, length=25,
                                      allow_unlabeled=True, sparse=False, return_indicator="dense",
                                      return_distributions=False, random_state=seed)
x_train, x_test, y_train, y_test = train_test_split(x, y, test_size=0.3, random_state=seed)
x_train, x_val, y_train, y_val = train_test_split(x_train, y_train, test_size=0.3, random_state=seed)
clf = MultiOutputClassifier(LogisticRegression()).fit(x_train, y_train)
proba = lambda xs: np.array(clf.predict_proba(xs))[:, :, 1].transpose()
out = {"y_val": y_val[:4000].astype(np.uint8), "y_proba_val": proba(x_val[:4000]).astype(np.float32),
       "y_test": y_test[:4000].astype(np.uint8), "y_proba_test": proba(x_test[:4000]).astype(np.float32)}
path = os.path.join(HERE, "reference_fixture.npz")
np.savez_compressed(path, **out)
print({k: v.shape for k, v in out.items()}, os.path.getsize(path) // 1024, "KiB")
