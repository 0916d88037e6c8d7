"""FactorizedEmbedding: tensorized drop-in for ``torch.nn.Embedding`` whose (num_embeddings, embedding_dim) table is stored in
factorized form and looked up WITHOUT being rebuilt (reference: tltorch/factorized_layers/factorized_embedding.py:10-236,
forward :96-117).  The lookup indexes the weight in factorized form (``weight[rows, :]``, torch_b200/embedding_ops.py)."""
import numpy as np
from ..utils.nvtx import annotate
import torch
from torch import nn

from ..factorized_tensors import TensorizedTensor, tensor_init
from ..utils import get_tensorized_shape


class FactorizedEmbedding(nn.Module):
    """Same constructor arguments as the reference: ``auto_tensorize`` picks the tensorized shapes with
    ``get_tensorized_shape``; otherwise ``tensorized_num_embeddings`` / ``tensorized_embedding_dim`` must multiply out to
    ``num_embeddings`` / ``embedding_dim``.  ``n_layers > 1`` parametrizes several tables with one jointly factorized tensor."""

    def __init__(self, num_embeddings, embedding_dim, auto_tensorize=True, n_tensorized_modes=3, tensorized_num_embeddings=None,
                 tensorized_embedding_dim=None, factorization="blocktt", rank=8, n_layers=1, device=None, dtype=None):
        super().__init__()
        if auto_tensorize:
            if tensorized_num_embeddings is not None and tensorized_embedding_dim is not None:
                raise ValueError("Either use auto_tensorize or specify tensorized_num_embeddings and tensorized_embedding_dim.")
            tensorized_num_embeddings, tensorized_embedding_dim = get_tensorized_shape(
                in_features=num_embeddings, out_features=embedding_dim, order=n_tensorized_modes, min_dim=2, verbose=False)
        else:
            computed_num_embeddings = np.prod(tensorized_num_embeddings)
            computed_embedding_dim = np.prod(tensorized_embedding_dim)
            if computed_num_embeddings != num_embeddings:
                raise ValueError("Tensorized embeddding number {} does not match num_embeddings argument {}".format(
                    computed_num_embeddings, num_embeddings))
            if computed_embedding_dim != embedding_dim:
                raise ValueError("Tensorized embeddding dimension {} does not match embedding_dim argument {}".format(
                    computed_embedding_dim, embedding_dim))
        self.num_embeddings = num_embeddings
        self.embedding_dim = embedding_dim
        self.tensor_shape = (tuple(tensorized_num_embeddings), tuple(tensorized_embedding_dim))
        self.weight_shape = (self.num_embeddings, self.embedding_dim)
        self.n_layers = n_layers
        if n_layers > 1:
            self.tensor_shape = (n_layers,) + self.tensor_shape
            self.weight_shape = (n_layers,) + self.weight_shape
        self.factorization = factorization
        self.weight = TensorizedTensor.new(self.tensor_shape, rank=rank, factorization=self.factorization, device=device, dtype=dtype)
        self.reset_parameters()
        self.rank = self.weight.rank

    def reset_parameters(self):
        # initialisation of TT-Rec (Yin et al.), as in the reference (:88-94)
        target_stddev = 1 / np.sqrt(3 * self.num_embeddings)
        with torch.no_grad():
            tensor_init(self.weight, std=target_stddev)

    @annotate(lambda m: "tltorch.FactorizedEmbedding")
    def forward(self, input, indices=0):
        output_shape = (*input.shape, self.embedding_dim)
        flattened_input = input.reshape(-1)
        if self.n_layers == 1:
            if indices != 0:
                raise ValueError(f"Only one embedding was parametrized (n_layers=1) but tried to access {indices}.")
            embeddings = self.weight[flattened_input, :]
        else:
            embeddings = self.weight[indices, flattened_input, :]
        if not torch.is_tensor(embeddings):          # CPTensorized stays factorized when indexed
            embeddings = embeddings.to_matrix()
        return embeddings.reshape(output_shape)

    @classmethod
    def from_embedding(cls, embedding_layer, rank=8, factorization="blocktt", n_tensorized_modes=2, decompose_weights=True,
                       auto_tensorize=True, decomposition_kwargs=dict(), **kwargs):
        num_embeddings, embedding_dim = embedding_layer.weight.shape
        instance = cls(num_embeddings, embedding_dim, auto_tensorize=auto_tensorize, factorization=factorization,
                       n_tensorized_modes=n_tensorized_modes, rank=rank, **kwargs)
        if decompose_weights:
            with torch.no_grad():
                instance.weight.init_from_matrix(embedding_layer.weight.data, **decomposition_kwargs)
        else:
            instance.reset_parameters()
        return instance

    @classmethod
    def from_embedding_list(cls, embedding_layer_list, rank=8, factorization="blocktt", n_tensorized_modes=2, decompose_weights=True,
                            auto_tensorize=True, decomposition_kwargs=dict(), **kwargs):
        n_layers = len(embedding_layer_list)
        num_embeddings, embedding_dim = embedding_layer_list[0].weight.shape
        for i, layer in enumerate(embedding_layer_list[1:]):
            new_num_embeddings, new_embedding_dim = layer.weight.shape
            if num_embeddings != new_num_embeddings:
                raise ValueError("All embedding layers must have the same num_embeddings."
                                 f"Yet, got embedding_layer_list[0] with num_embeddings={num_embeddings} "
                                 f" and embedding_layer_list[{i+1}] with num_embeddings={new_num_embeddings}.")
            if embedding_dim != new_embedding_dim:
                raise ValueError("All embedding layers must have the same embedding_dim."
                                 f"Yet, got embedding_layer_list[0] with embedding_dim={embedding_dim} "
                                 f" and embedding_layer_list[{i+1}] with embedding_dim={new_embedding_dim}.")
        instance = cls(num_embeddings, embedding_dim, n_tensorized_modes=n_tensorized_modes, auto_tensorize=auto_tensorize,
                       factorization=factorization, rank=rank, n_layers=n_layers, **kwargs)
        if decompose_weights:
            weight_tensor = torch.stack([layer.weight.data for layer in embedding_layer_list])
            with torch.no_grad():
                instance.weight.init_from_matrix(weight_tensor, **decomposition_kwargs)
        else:
            instance.reset_parameters()
        return instance

    def get_embedding(self, indices):
        if self.n_layers == 1:
            raise ValueError("A single linear is parametrized, directly use the main class.")
        return SubFactorizedEmbedding(self, indices)


class SubFactorizedEmbedding(nn.Module):
    """One of the embeddings of a jointly factorized FactorizedEmbedding (shares the parent's parameters)."""

    def __init__(self, main_layer, indices):
        super().__init__()
        self.main_layer = main_layer
        self.indices = indices

    def forward(self, x):
        return self.main_layer(x, self.indices)

    def extra_repr(self):
        return ""
