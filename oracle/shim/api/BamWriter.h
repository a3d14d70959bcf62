// TEST INFRASTRUCTURE: the reference includes this header but uses nothing from it on any path.
