"""TEST-ONLY stand-in for `diffusers==0.13.1` (absent from this image, no network).

Restates — from the published behaviour of that version, SURVEY.md Appendix C — exactly the classes
the reference's model/*.py import, so that /root/reference/model/{attention,unet_2d_blocks,
unet_2d_condition,pipeline}.py run UNMODIFIED as the parity anchor for oracle/.  Never imported by
the product package.
"""


class AutoencoderKL:  # only used as a type annotation by model/pipeline.py:16
    pass
