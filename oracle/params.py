"""Oracle helper (TEST INFRASTRUCTURE ONLY): random proposal-network parameters under the reference's state_dict names,
built on the CPU with the reference's own initialisers (nn.Linear / nn.LSTM defaults, normal_() embeddings;
pyprob/nn/embedding_feedforward.py:8-33, inference_network_lstm.py:29-72) — what the CPU baseline legs of bench.py and the
CPU tests feed to oracle.network without touching the CUDA library."""
import torch
import torch.nn as nn


def _ff_dims(in_dim, out_dim, depth):
    if depth == 1:
        return [(in_dim, out_dim)]
    hidden = int((in_dim + out_dim) / 2)
    return [(in_dim, hidden)] + [(hidden, hidden)] * (depth - 2) + [(hidden, out_dim)]


def random_params(observe, addresses, lstm_dim=512, K=10, sample_dim=4, addr_dim=64, type_dim=8, seed=0):
    """observe: list of (name, in_dim, out_dim, depth); addresses: list of (address, family, num_categories)."""
    torch.manual_seed(seed)
    P = {}

    def lin(prefix, i, o):
        m = nn.Linear(i, o)
        P[prefix + '.weight'], P[prefix + '.bias'] = m.weight.detach().clone(), m.bias.detach().clone()

    E = 0
    for name, in_dim, out_dim, depth in observe:
        for l, (a, b) in enumerate(_ff_dims(in_dim, out_dim, depth)):
            lin('_layers_observe_embedding.{}._layers.{}'.format(name, l), a, b)
        E += out_dim
    for l, (a, b) in enumerate(_ff_dims(E, E, 2)):
        lin('_layers_observe_embedding_final._layers.{}'.format(l), a, b)
    I = E + sample_dim + 2 * (addr_dim + type_dim)
    lstm = nn.LSTM(I, lstm_dim, 1)
    for k in ('weight_ih_l0', 'weight_hh_l0', 'bias_ih_l0', 'bias_hh_l0'):
        P['_layers_lstm.' + k] = getattr(lstm, k).detach().clone()
    for address, family, C in addresses:
        P['_layers_address_embedding.' + address] = torch.randn(addr_dim)
        if '_layers_distribution_type_embedding.' + family not in P:
            P['_layers_distribution_type_embedding.' + family] = torch.randn(type_dim)
        out = C if family == 'Categorical' else 3 * K
        lin('_layers_sample_embedding.{}._layers.0'.format(address), C if family == 'Categorical' else 1, sample_dim)
        lin('_layers_proposal.{}._ff._layers.0'.format(address), lstm_dim, int((lstm_dim + out) / 2))
        lin('_layers_proposal.{}._ff._layers.1'.format(address), int((lstm_dim + out) / 2), out)
    return P
