/* placeholder translation unit: filled by the counting oracle (declare / intersect / stats) */
