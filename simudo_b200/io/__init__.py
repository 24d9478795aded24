from .output_writer import (OutputWriter, MetaExtractorBandInfo, MetaExtractorIntegrals)
