"""reveal_b200 -- B200-native index build, MUM sweeps and recursion steps behind the `reveallib`
extension surface of jasperlinthorst/reveal (`reveal_b200.reveallib`, `reveallib64`), and the REM
driver on top of them (`reveal_b200.rem`).  See DESIGN.md / INTEGRATION.md."""
__version__ = "0.1"
