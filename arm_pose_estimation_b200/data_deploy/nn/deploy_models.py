"""Hashes of the three deployed MC-dropout LSTMs (``data_deploy/nn/deploy_models.py:4-7`` of the reference)."""
from enum import Enum


class LSTM(Enum):
    WATCH_PHONE_POCKET = "670b66fa7664252d1cfb3b5a8a362002ffeeba5c"
    WATCH_PHONE_UARM = "7cb5cdf94ef4c66388c7f15f642005d5e008146a"
    WATCH_ONLY = "04f4ad63bfccb3668f7598c9375403e10b1fae2a"
