import csv,sys,subprocess
for f in sys.argv[1:]:
    out=subprocess.run(['ncu','-i',f,'--page','raw','--csv'],capture_output=True,text=True).stdout
    rows=list(csv.reader(out.splitlines())); hdr=rows[0]; units=rows[1]
    want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_sectors_srcunit_ltcfabric.sum','lts__t_sectors_srcunit_ltcfabric_lookup_miss.sum',
          'lts__t_sectors_srcunit_ltcfabric_lookup_hit.sum','lts__t_sectors_aperture_peer.sum','lts__t_sectors_aperture_peer_op_read.sum','lts__t_sectors_aperture_peer_op_write.sum',
          'lts__t_sectors_srcunit_tex_aperture_peer.sum','l1tex__m_xbar2l1tex_read_sectors_mem_lg_op_ld.sum','nvlink__bandwidth','nvlink__is_nvswitch_connected','nvlink__count_physical',
          'smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','sm__warps_active.avg.pct_of_peak_sustained_active',
          'lts__t_bytes.sum','lts__t_sectors_op_red.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum','l1tex__t_requests_pipe_lsu_mem_global_op_red.sum']
    for r in rows[2:]:
        print('==',f.split('/')[-1],r[hdr.index('Kernel Name')],'grid',r[hdr.index('Grid Size')])
        for k in hdr:
            if k in want or ('peer' in k and k.endswith('.sum')) or k.startswith('nvl'):
                v=r[hdr.index(k)]
                if v not in ('','0','n/a'): print('   %-80s %-10s %s'%(k,units[hdr.index(k)],v))
