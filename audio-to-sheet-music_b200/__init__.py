"""B200-native (sm_100a) segment-batched separation forward of the text-conditioned HTDemucs
(``AudioTextHTDemucs``) of savage-hacker14/audio-to-sheet-music.  Import as ``athtd_b200``."""
from .lib import AthtdError, LIB_PATH, load as load_library          # noqa: F401
from .model import AudioTextHTDemucsB200, HTDemucsParams, TextCrossAttention, FreqDecoder, TimeDecoder  # noqa: F401
from .separation import (B200SeparationModel, SeparationModel, STEMS, segment_plan, OlaTables, gather_chunks,  # noqa: F401
                         chunk_ola)
from .engine import Engine, Plan                                     # noqa: F401
from . import distributed                                     # noqa: F401
from . import metrics                                         # noqa: F401
from . import audio                                           # noqa: F401
from .audio import DeviceResampler, prepare_mixture           # noqa: F401
from .clap_text import ClapTextModelWithProjectionB200, ClapModelTextB200   # noqa: F401
