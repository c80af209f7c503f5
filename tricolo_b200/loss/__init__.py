from .nt_xent import NTXentLoss, calculate_losses, trimodal_ntxent  # noqa: F401
