namespace boost { struct null_deleter { template <class T> void operator()(T*) const {} }; }
