"""firedrake_ts_b200 -- B200-native assembly engine behind the firedrake-ts API."""
