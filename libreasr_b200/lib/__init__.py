"""Mirror of the reference's ``libreasr.lib`` package for the inference path."""
from .models import Transducer, Encoder, Predictor, Joint  # noqa: F401
