"""`GraphKernel`: the name -> kernel dispatcher of the reference
(grakel/graph_kernels.py:77-571), restricted to the kernels of the device hot path:

    "weisfeiler_lehman" / "WL"  (framework; base "vertex_histogram"/"subtree_wl"/"VH"/"ST-WL")
    "vertex_histogram" / "subtree_wl" / "VH" / "ST-WL"
    "shortest_path" / "SP"      (+ with_labels, algorithm_type, as_attributes)
    "edge_histogram" / "EH"
    "core_framework" / "CORE"   (framework; base "shortest_path" or "weisfeiler_lehman" ...)
    "weisfeiler_lehman_optimal_assignment" / "WL-OA"

Any other reference kernel name raises NotImplementedError (not ValueError, which
the reference reserves for unknown names).  The Nystroem option is host-side
post-processing of the kernel matrix and is kept (graph_kernels.py:311-335, 367, 403).
"""
import copy
import warnings

import numpy as np
from scipy.linalg import svd
from sklearn.base import BaseEstimator, TransformerMixin
from sklearn.utils import check_random_state
from sklearn.utils.validation import check_is_fitted

from .core_framework import CoreFramework
from .kernels import (EdgeHistogram, ShortestPath, ShortestPathAttr, VertexHistogram, WeisfeilerLehman,
                      WeisfeilerLehmanOptimalAssignment)

_VH = ("vertex_histogram", "subtree_wl", "VH", "ST-WL")
_SP = ("shortest_path", "SP")
_WL = ("weisfeiler_lehman", "WL")
_EH = ("edge_histogram", "EH")
_CORE = ("core_framework", "CORE")
_WLOA = ("weisfeiler_lehman_optimal_assignment", "WL-OA")
# names the reference knows but that are outside the hot path (graph_kernels.py:38-64)
_OTHER = {"random_walk", "RW", "graphlet_sampling", "GR", "subgraph_matching", "SM",
          "multiscale_laplacian", "ML", "lovasz_theta", "LOVT", "svm_theta", "SVMT", "neighborhood_hash", "NH",
          "neighborhood_subgraph_pairwise_distance", "NSPD", "odd_sth", "ODD", "propagation", "PR",
          "pyramid_match", "PM", "graph_hopper", "GH", "hadamard_code", "HC"}
default_n_components = 100


class GraphKernel(BaseEstimator, TransformerMixin):
    def __init__(self, kernel="shortest_path", normalize=False, verbose=False, n_jobs=None, random_state=None,
                 Nystroem=False):
        self.kernel = kernel
        self.normalize = normalize
        self.verbose = verbose
        self.n_jobs = n_jobs
        self.random_state = random_state
        self.Nystroem = Nystroem
        self._initialized = {"kernel": False, "Nystroem": False, "random_state": False, "normalize": False,
                             "verbose": False, "n_jobs": False}

    def fit(self, X, y=None):
        self.initialize()
        if bool(self.nystroem_):
            X = list(X)
            nx = len(X)
            if self.nystroem_ > nx:
                n_components = nx
                warnings.warn("n_components > n_samples. This is not possible.\nn_components was set to n_samples, "
                              "which results in inefficient evaluation of the full kernel.")
            else:
                n_components = self.nystroem_
            n_components = min(nx, n_components)
            inds = self.random_state_.permutation(nx)
            basis = [X[i] for i in inds[:n_components]]
            U, S, V = svd(self.kernel_.fit_transform(basis))
            S = np.maximum(S, 1e-12)
            self.nystroem_ = n_components
            self.nystroem_normalization_ = np.dot(U / np.sqrt(S), V)
            self.components_ = basis
            self.component_indices_ = inds
        else:
            self.kernel_.fit(X)
        return self

    def transform(self, X):
        check_is_fitted(self, "kernel_")
        if hasattr(self, "nystroem_") and bool(self.nystroem_):
            check_is_fitted(self, "components_")
            return self.kernel_.transform(X).dot(self.nystroem_normalization_.T)
        return self.kernel_.transform(X)

    def fit_transform(self, X, y=None):
        self.initialize()
        if bool(self.nystroem_):
            self.fit(X)
            return self.kernel_.transform(X).dot(self.nystroem_normalization_.T)
        return self.kernel_.fit_transform(X)

    def initialize(self):
        if not self._initialized["Nystroem"]:
            if type(self.Nystroem) not in [int, bool]:
                raise ValueError("Nystroem parameter must be an int, indicating the number of components or a boolean")
            elif self.Nystroem is False:
                self.nystroem_ = False
            elif self.Nystroem in [0, -1] or self.Nystroem is True:
                self.nystroem_ = default_n_components
            elif self.Nystroem <= 0:
                raise ValueError("number of nystroem components must be positive")
            else:
                self.nystroem_ = self.Nystroem
            self._initialized["Nystroem"] = True
        if any(not self._initialized[p] for p in ["random_state", "normalize", "verbose", "n_jobs", "kernel"]):
            if not self._initialized["random_state"]:
                self.random_state_ = check_random_state(self.random_state)
            k = self.kernel
            if type(k) is dict or type(k) is str:
                k = [self.kernel]
            elif type(k) is not list:
                raise ValueError('A "kernel" must be defined at the __init__ function of the graph kernel generic '
                                 "wrapper. Valid kernel types are dict, str, and list of dict or str.")
            hidden = {"verbose": self.verbose, "normalize": self.normalize, "n_jobs": self.n_jobs}
            cls, params = self.make_kernel_(copy.deepcopy(k), hidden)
            self.kernel_ = cls(**params)
            for p in ["random_state", "normalize", "verbose", "n_jobs", "kernel"]:
                self._initialized[p] = True

    def make_kernel_(self, kernel_list, hidden_args):
        kernel = kernel_list.pop(0)
        if type(kernel) is str:
            name, kernel = str(kernel), dict()
        elif type(kernel) is not dict:
            raise ValueError("each element of the list of kernels must be a dictionary or a string")
        else:
            if "name" not in kernel:
                raise ValueError('each dictionary concerning a kernel must have a "name" parameter designating '
                                 "the kernel")
            name = kernel.pop("name")
        for key, val in hidden_args.items():
            if key in kernel:
                warnings.warn("Overriding global kernel attribute " + str(key) + " with " + str(val) +
                              ". Please set this attribute as an argument of GraphKernel.")
            kernel[key] = val
        if name in _VH or name in _SP or name in _EH or name in _WLOA:
            if len(kernel_list) != 0:
                warnings.warn("Kernel List not empty while reaching a base-kernel - the rest kernel names will be "
                              "ignored")
            if name in _VH:
                return VertexHistogram, kernel
            if name in _EH:
                return EdgeHistogram, kernel
            if name in _WLOA:  # graph_kernels.py:540-541
                return WeisfeilerLehmanOptimalAssignment, kernel
            if kernel.pop("as_attributes", False):
                return ShortestPathAttr, kernel
            return ShortestPath, kernel
        if name in _WL:
            if len(kernel_list):
                kernel["base_graph_kernel"] = self.make_kernel_(kernel_list, {})
            return WeisfeilerLehman, kernel
        if name in _CORE:  # graph_kernels.py:527-531
            if len(kernel_list):
                kernel["base_graph_kernel"] = self.make_kernel_(kernel_list, {})
            return CoreFramework, kernel
        if name in _OTHER:
            raise NotImplementedError("kernel '" + str(name) + "' is outside the device hot path of grakel_b200 "
                                      "(WL-subtree, vertex histogram, shortest path)")
        raise ValueError("Unsupported kernel: " + str(name))

    def set_params(self, **params):
        params = copy.deepcopy(params)
        for key in params:
            key, delim, sub_key = key.partition("__")
            if delim:
                if sub_key in self._initialized:
                    self._initialized[sub_key] = False
            elif key in self._initialized:
                self._initialized[key] = False
        super().set_params(**params)
        return self
