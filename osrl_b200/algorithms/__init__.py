from .bc import BC, BCTrainer  # noqa: F401
from .bcql import BCQL, BCQLTrainer  # noqa: F401
from .cpq import CPQ, CPQTrainer  # noqa: F401
from .bearl import BEARL, BEARLTrainer  # noqa: F401
from .cdt import CDT, CDTTrainer  # noqa: F401
from .coptidice import COptiDICE, COptiDICETrainer  # noqa: F401
