"""Derive the small data fixtures from the reference's data files (run in the dev container, where /root/reference
is mounted).  The GPU box has no /root/reference, so tests and bench read these fixtures instead.

    python tests/golden/make_fixtures.py

bikeshare.npz   config C2 (main.cpp:54-71): X = trainP7_1.csv with column 1 erased (main.cpp:66-67) -> 4337 x 10,
                Y = trainYP2_1.csv -> 4337 x 1; validate / predict sets likewise.
mnist_labels.npy config C1 (main.cpp:113-134): trainY.csv ++ validateY.csv -> 42000 class ids (the pixel CSVs are
                missing from the reference checkout; pixels are synthesised by the tests / bench).
"""
import os
import numpy as np

REF = os.environ.get("B2NN_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def load(name):
    return np.loadtxt(os.path.join(REF, name), delimiter=",", dtype=np.float64, ndmin=2).astype(np.float32)


def drop_col1(a):
    return np.delete(a, 1, axis=1)


def main():
    np.savez_compressed(
        os.path.join(HERE, "bikeshare.npz"),
        x=drop_col1(load("bikeshare/trainP7_1.csv")), y=load("bikeshare/trainYP2_1.csv"),
        xv=drop_col1(load("bikeshare/validateP7_1.csv")), yv=load("bikeshare/validateY1.csv"),
        xp=drop_col1(load("bikeshare/testPF1_1.csv")))
    lab = np.concatenate([load("MNIST/trainY.csv")[:, 0], load("MNIST/validateY.csv")[:, 0]]).astype(np.uint8)
    np.save(os.path.join(HERE, "mnist_labels.npy"), lab)
    print("bikeshare", {k: v.shape for k, v in np.load(os.path.join(HERE, "bikeshare.npz")).items()})
    print("mnist labels", lab.shape, np.bincount(lab))


if __name__ == "__main__":
    main()
