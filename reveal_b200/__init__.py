"""reveal_b200 -- B200-native index build + MUM sweeps behind the `reveallib`
extension surface of jasperlinthorst/reveal.  See DESIGN.md / INTEGRATION.md."""
__version__ = "0.1"
