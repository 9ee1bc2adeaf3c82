"""Loader factory under the reference's module path (train.py:122 calls
`imitation_dataset.sequential_train_val_test_iterator(hparams)`). The implementation is the
device-staged sequential pipeline in carla_imitation_learning_b200.data."""
from carla_imitation_learning_b200.data import (  # noqa: F401
    SequentialFrames, continous_to_discreet, sequential_train_val_test_iterator, synthetic_sequence)
