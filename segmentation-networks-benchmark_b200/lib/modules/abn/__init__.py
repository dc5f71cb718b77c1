from .bn import InPlaceABN  # noqa: F401
from .functions import inplace_abn  # noqa: F401
