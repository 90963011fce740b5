"""Pickle helpers with the reference's names, formats and prints (utils.py:242-258) and the half-moons data set
(utils.py:67-92, 209-236), the one data set the reference builds itself (MNIST / Fashion-MNIST come from keras
downloads and stay out of scope)."""
import os
import pickle as pkl
import random

import numpy as np


def save_to_pickle(data, path, filename):
    print("\nSaving pickle: ", path + filename)
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path + filename, 'wb') as f:
        pkl.dump(data, f)


def load_from_pickle(path):
    print("\nLoading from pickle: ", path)
    with open(path, 'rb') as f:
        u = pkl._Unpickler(f)
        u.encoding = 'latin1'
        data = u.load()
    return data


def load_half_moons(channels="first", n_samples=30000):
    """sklearn make_moons(30000, noise 0.1, random_state 0), min-max normalised jointly over both coordinates, last
    20 % as the test set, inputs shaped [N, 1, 2, 1] ("image-like"), one-hot labels with 2 classes (utils.py:67-92)."""
    from sklearn.datasets import make_moons
    x, y = make_moons(n_samples=n_samples, shuffle=True, noise=0.1, random_state=0)
    x, y = (x.astype('float32'), y.astype('float32'))
    x = (x - np.min(x)) / (np.max(x) - np.min(x))
    split_size = int(0.8 * len(x))
    x_train, y_train = x[:split_size], y[:split_size]
    x_test, y_test = x[split_size:], y[split_size:]
    n_channels, n_coords = 1, 2
    if channels == "first":
        x_train = x_train.reshape(x_train.shape[0], n_channels, n_coords, 1)
        x_test = x_test.reshape(x_test.shape[0], n_channels, n_coords, 1)
    elif channels == "last":
        x_train = x_train.reshape(x_train.shape[0], 1, n_coords, n_channels)
        x_test = x_test.reshape(x_test.shape[0], 1, n_coords, n_channels)
    input_shape = x_train.shape[1:]
    num_classes = 2
    y_train = np.eye(num_classes, dtype='float32')[y_train.astype('int64')]      # keras.utils.to_categorical
    y_test = np.eye(num_classes, dtype='float32')[y_test.astype('int64')]
    return x_train, y_train, x_test, y_test, input_shape, num_classes


def load_dataset(dataset_name, n_inputs=None, channels="first", shuffle=False):
    """utils.py:209-236 for the data set that needs no download."""
    if dataset_name != "half_moons":
        raise AssertionError("\nDataset not available.")
    x_train, y_train, x_test, y_test, input_shape, num_classes = load_half_moons()
    if n_inputs:
        x_train, y_train, x_test, y_test = (x_train[:n_inputs], y_train[:n_inputs], x_test[:n_inputs], y_test[:n_inputs])
    print('x_train shape =', x_train.shape, '\nx_test shape =', x_test.shape)
    print('y_train shape =', y_train.shape, '\ny_test shape =', y_test.shape)
    if shuffle is True:
        random.seed(0)
        idxs = np.random.permutation(len(x_train))
        x_train, y_train = (x_train[idxs], y_train[idxs])
        idxs = np.random.permutation(len(x_test))
        x_test, y_test = (x_test[idxs], y_test[idxs])
    return x_train, y_train, x_test, y_test, input_shape, num_classes


def data_loaders(dataset_name, batch_size, n_inputs, channels="first", shuffle=True):
    """utils.py:24-37."""
    from torch.utils.data import DataLoader
    x_train, y_train, x_test, y_test, input_shape, num_classes = \
        load_dataset(dataset_name=dataset_name, n_inputs=n_inputs, channels=channels, shuffle=shuffle)
    train_loader = DataLoader(dataset=list(zip(x_train, y_train)), batch_size=batch_size, shuffle=shuffle,
                              worker_init_fn=np.random.seed(0), num_workers=0)
    test_loader = DataLoader(dataset=list(zip(x_test, y_test)), batch_size=batch_size, shuffle=shuffle,
                             worker_init_fn=np.random.seed(0), num_workers=0)
    return train_loader, test_loader, input_shape, num_classes
