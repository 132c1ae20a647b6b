"""gcc_b200: B200-native GCC cooperative-compression training step."""
