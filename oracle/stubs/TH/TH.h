// empty stub: the reference file includes this header but uses nothing from it
