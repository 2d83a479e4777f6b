"""Drop-in mirror of the reference's ``models/`` files for the routed interaction stack.

Same class names, constructor arguments, ``forward`` signatures, parameter names / shapes / creation
order (so ``state_dict`` round-trips and the same seed gives the same initial weights) as
``models/{InteractionModule,DynamicInteraction,Cells,Router,SelfAttention,Refinement,XModules}.py``;
the arithmetic runs in ``libd2r_b200.so``.  See INTEGRATION.md for the two-line shim that makes the
reference's ``modeling_unimo.py`` import these instead of its own.
"""
from .InteractionModule import InteractionModule, Reversed_InteractionModule, run_pair  # noqa: F401
