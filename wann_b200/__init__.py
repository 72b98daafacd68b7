"""wann_b200: B200-native (sm_100a) batched policy-net leaf evaluation for WaNN.

Only the hot path: board encode -> conv stack (tcgen05) -> FC-155 -> softmax ->
legal-move gather/sort, behind the reference's `policy_net.evaluate` / `sess.run` surface."""
from ._lib import WannError, LIB_PATH  # noqa: F401
from .evaluator import PolicyEvaluator  # noqa: F401
from .multi import MultiGpuEvaluator  # noqa: F401
from .forest import SearchForest  # noqa: F401
from .session import (NeuralNetsCombined_128, Session, call_policy_net, instantiate_session,  # noqa: F401
                      instantiate_session_both, instantiate_session_both_128)
from . import boards, weights  # noqa: F401
