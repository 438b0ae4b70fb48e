from .buffer import Basic_Buffer_Type, Test_Buffer_Type, Train_Buffer_Type, BITS_SET, CommBuffer  # noqa: F401
from .comm import Communicator  # noqa: F401
