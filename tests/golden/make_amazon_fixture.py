"""Writes tests/golden/amazon_hpmn_sample.pkl: the first 256 train / 128 test tuples of the reference's sample
dataset /root/reference/data/amazon/dataset_hpmn.pkl (BASELINE.json configs[0]; tuple layout
`(label, item_part [100][3], item_len, user_part [100][2], user_len)`, code/util.py:152-159), re-pickled in the same
three-object layout `train list, test list, feature_size` that code/hpmn.py:571-575 reads, binary protocol 2 to
keep the file small.  Run in the build container (the GPU box has no /root/reference):

    python tests/golden/make_amazon_fixture.py
"""
import os
import pickle

SRC = "/root/reference/data/amazon/dataset_hpmn.pkl"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "amazon_hpmn_sample.pkl")
N_TRAIN, N_TEST = 256, 128


def main():
    with open(SRC, "rb") as f:
        train = pickle.load(f, encoding="latin1")
        test = pickle.load(f, encoding="latin1")
        feature_size = pickle.load(f, encoding="latin1")
    with open(DST, "wb") as f:
        pickle.dump(train[:N_TRAIN], f, protocol=2)
        pickle.dump(test[:N_TEST], f, protocol=2)
        pickle.dump(int(feature_size), f, protocol=2)
    print(DST, os.path.getsize(DST), "bytes; feature_size", feature_size)


if __name__ == "__main__":
    main()
