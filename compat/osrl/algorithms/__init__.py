"""osrl.algorithms -> osrl_b200.algorithms (reference: osrl/algorithms/__init__.py)."""
from osrl_b200.algorithms import *  # noqa: F401,F403
from osrl_b200.algorithms import (BC, BCQL, BEARL, CDT, CPQ, BCQLTrainer, BCTrainer, BEARLTrainer, CDTTrainer,  # noqa: F401
                                  COptiDICE, COptiDICETrainer, CPQTrainer)
