/* empty stub: the headless reference harness never touches OpenGL (see oracle/Makefile) */
