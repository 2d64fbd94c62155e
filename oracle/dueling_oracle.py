"""TEST INFRASTRUCTURE ONLY (imported by tests/, never by the product).

CPU restatement (numpy float32) of the DQN head blocks the reference ships:
    DuelingArchitecture.forward                  drl/agents/architectures/stateless/dueling.py:50-57
      with MLP (drl/agents/architectures/stateless/mlp.py:32-41): Linear -> ReLU -> Linear -> ReLU, i.e. the
      advantage and value streams END in a ReLU (the reference's list comprehension appends one after every Linear)
    EpsilonGreedyCategoricalPolicyHead.forward   drl/agents/heads/policy_heads.py:197-226
Pinned to the unmodified reference by tests/golden/dueling.npz (oracle/make_golden.py::gen_dueling); the reference's
own test (test_dueling.py:7-21) only checks all-zero weights, which the golden file also contains."""
import numpy as np

F32 = np.float32

PARAM_KEYS = ['_proj.0.weight', '_proj.0.bias',
              '_advantage_stream._network.0.weight', '_advantage_stream._network.0.bias',
              '_advantage_stream._network.2.weight', '_advantage_stream._network.2.bias',
              '_value_stream._network.0.weight', '_value_stream._network.0.bias',
              '_value_stream._network.2.weight', '_value_stream._network.2.bias']


def _lin(x, w, b):
    return (x.astype(F32) @ w.astype(F32).T + b.astype(F32)).astype(F32)


def dueling_forward(params, features):
    """params: dict keyed like the reference module's state_dict (PARAM_KEYS); features [nb, input_dim]."""
    relu = lambda t: np.maximum(t, F32(0))
    proj = relu(_lin(features, params[PARAM_KEYS[0]], params[PARAM_KEYS[1]]))                 # dueling.py:51
    adv = relu(_lin(relu(_lin(proj, params[PARAM_KEYS[2]], params[PARAM_KEYS[3]])), params[PARAM_KEYS[4]], params[PARAM_KEYS[5]]))
    adv = (adv - adv.mean(axis=-1, keepdims=True, dtype=F32)).astype(F32)                     # :53
    val = relu(_lin(relu(_lin(proj, params[PARAM_KEYS[6]], params[PARAM_KEYS[7]])), params[PARAM_KEYS[8]], params[PARAM_KEYS[9]]))
    return (adv + val).astype(F32)                                                            # :55


def epsilon_greedy_probs(q_values, epsilon):
    """(1 - eps) * one_hot(argmax q) + eps * uniform, before Categorical's own normalisation (policy_heads.py:219-224)."""
    q = q_values.astype(F32)
    nb, A = q.shape
    greedy = np.zeros_like(q)
    greedy[np.arange(nb), np.argmax(q, axis=-1)] = F32(1)                                     # first maximum, like torch.argmax
    uniform = (np.ones_like(q) / np.ones_like(q).sum(axis=-1, keepdims=True)).astype(F32)
    return (F32(1 - epsilon) * greedy + F32(epsilon) * uniform).astype(F32)
