from .nt_xent import NTXentLoss, calculate_losses, trimodal_ntxent, trimodal_ntxent_total  # noqa: F401
from .triplet import TripletLoss  # noqa: F401,E402
